def get_cmap(name=None):
    return None
