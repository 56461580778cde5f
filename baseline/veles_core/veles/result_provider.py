from zope.interface import Interface


class IResultProvider(Interface):
    pass
