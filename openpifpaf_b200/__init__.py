"""openpifpaf_b200 -- B200-native inference hot path for OpenPifPaf (backbone + CIF/CAF heads + CifCaf decode).

The package name starts with ``openpifpaf_`` on purpose: the reference discovers plugins by that prefix
(openpifpaf/plugin.py:17-40) and calls ``register()``.  Importing this package never imports the reference.
"""
__version__ = '0.1.0'


def register():
    """Plugin hook of the reference (openpifpaf/plugin.py:36-40): adds the CifCafB200 decoder to
    ``openpifpaf.DECODERS`` with a higher priority than the reference's CPU CifCaf."""
    from . import plugin
    plugin.register()
