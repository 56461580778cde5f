"""``veles.result_provider.IResultProvider``: units that publish final metrics."""


class IResultProvider(object):
    __required__ = ("get_metric_names", "get_metric_values")
