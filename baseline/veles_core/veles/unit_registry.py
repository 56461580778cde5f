from veles.mapped_object_registry import MappedObjectsRegistry


class UnitRegistry(type):
    enabled = True
    units = set()
    hidden_units = set()

    def __init__(cls, name, bases, clsdict):
        if clsdict.get("hide_from_registry", False):
            UnitRegistry.hidden_units.add(cls)
        else:
            UnitRegistry.units.add(cls)
        super(UnitRegistry, cls).__init__(name, bases, clsdict)


class MappedUnitRegistry(UnitRegistry, MappedObjectsRegistry):
    pass
