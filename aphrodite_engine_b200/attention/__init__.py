from .paged_attn import PagedAttention  # noqa: F401
