class VelesException(Exception):
    pass


class BadFormatError(VelesException):
    pass


class AlreadyExistsError(VelesException):
    pass


class NotExistsError(VelesException):
    pass


class Bug(VelesException):
    pass


class MasterSlaveCommunicationError(VelesException):
    pass
