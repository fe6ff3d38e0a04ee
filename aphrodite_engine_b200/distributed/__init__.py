from .custom_all_reduce import CustomAllreduce  # noqa: F401
