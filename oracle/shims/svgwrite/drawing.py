class Drawing:
    pass
