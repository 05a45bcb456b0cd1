__version__ = '0.1.0'
# Mitty release whose read-generation hot path this engine mirrors (mitty/version.py:1)
__mitty_version__ = '2.7.3.dev0'
