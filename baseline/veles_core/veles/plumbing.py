from veles.workflow import FireStarter, Repeater, StartPoint, EndPoint  # noqa: F401
