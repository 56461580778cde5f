from zope.interface import Interface


class IDistributable(Interface):
    pass


class Distributable(object):
    def __init__(self, *args, **kwargs):
        super(Distributable, self).__init__(*args, **kwargs)


class TriviallyDistributable(object):
    def generate_data_for_master(self):
        return None

    def generate_data_for_slave(self, slave):
        return None

    def apply_data_from_master(self, data):
        pass

    def apply_data_from_slave(self, data, slave):
        pass

    def drop_slave(self, slave):
        pass
