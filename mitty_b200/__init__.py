"""mitty_b200: a B200-native engine for Mitty's read-generation hot path
(generate-reads / corrupt-reads).  See DESIGN.md."""
from mitty_b200.version import __version__  # noqa: F401
