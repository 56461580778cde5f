from veles.workflow import DummyLauncher, DummyWorkflow, DummyUnit  # noqa: F401
