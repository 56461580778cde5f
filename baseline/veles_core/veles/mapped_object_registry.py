class MappedObjectsRegistry(type):
    """Metaclass: ``mapping = "evaluators"`` on the registry class creates the dict
    ``Registry.evaluators``; classes with ``MAPPING = "name"`` register under it."""
    mapping = "objects"
    base = object

    def __init__(cls, name, bases, clsdict):
        yours = set(cls.mro())
        mine = set(cls.base.mro()) if isinstance(cls.base, type) else set()
        left = yours - mine
        mapping = clsdict.get("MAPPING", None)
        reg = type(cls)
        table_name = reg.mapping
        table = reg.__dict__.get(table_name)
        if table is None:
            # the table lives on the class that declared ``mapping``
            owner = next((k for k in reg.__mro__ if "mapping" in k.__dict__), reg)
            table = owner.__dict__.get(table_name)
            if table is None:
                table = {}
                setattr(owner, table_name, table)
        if mapping and isinstance(mapping, str) and left:
            table[mapping] = cls
        super(MappedObjectsRegistry, cls).__init__(name, bases, clsdict)
