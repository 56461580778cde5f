"""The two names the reference uses from zope.interface (absent offline)."""


class Interface(object):
    pass


def implementer(*interfaces):
    def deco(cls):
        prev = getattr(cls, "__implemented__", ())
        cls.__implemented__ = tuple(prev) + tuple(interfaces)
        return cls
    return deco


def implementedBy(cls):
    return getattr(cls, "__implemented__", ())


def providedBy(obj):
    return getattr(type(obj), "__implemented__", ())
