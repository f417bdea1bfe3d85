class Rectangle:  # noqa: D401 - placeholder, never instantiated on the hot path
    pass
