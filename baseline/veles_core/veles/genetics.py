class Range(object):
    def __init__(self, default, *bounds):
        self.default = default
        self.bounds = bounds


def fix_config(cfg):
    for k, v in list(cfg.__dict__.items()):
        if isinstance(v, Range):
            setattr(cfg, k, v.default)
        elif hasattr(v, "__content__"):
            fix_config(v)
