from zope.interface import Interface


class IImageLoader(Interface):
    pass


class ImageLoader(object):
    """Marker base for image loaders (``isinstance(loader, ImageLoader)`` checks)."""
    pass
