"""abyss_b200: B200-native Bloom-filter de Bruijn graph unitig stage (abyss-bloom-dbg path)."""
__version__ = "0.1.0"
