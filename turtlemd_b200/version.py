"""Version: tracks the TurtleMD release whose API this package mirrors."""
__version__ = "2023.3.1+b200.1"
