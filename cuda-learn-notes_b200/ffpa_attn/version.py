__version__ = "0.0.2+b200"
