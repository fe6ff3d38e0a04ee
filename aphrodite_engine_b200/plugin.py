"""`aphrodite.general_plugins` entry point (reference: aphrodite/plugins/__init__.py:8-31): loads this
package's `_C` registrations into every engine / worker process. See INTEGRATION.md, option B."""


def register() -> None:
    from . import _native
    _native.load_torch_ops()
