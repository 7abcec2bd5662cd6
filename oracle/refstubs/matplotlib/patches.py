class Rectangle:  # noqa: D101
    pass
