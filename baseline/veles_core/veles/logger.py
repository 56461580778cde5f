import logging
import sys

_configured = False


def setup_logging(level=logging.WARNING):
    global _configured
    if not _configured:
        logging.basicConfig(level=level, stream=sys.stderr,
                            format="%(asctime)s %(levelname).1s %(name)s: %(message)s",
                            datefmt="%H:%M:%S")
        _configured = True
    logging.getLogger().setLevel(level)


class Logger(object):
    def __init__(self, *args, **kwargs):
        super(Logger, self).__init__()

    @property
    def logger(self):
        lg = self.__dict__.get("logger_")
        if lg is None:
            lg = logging.getLogger(type(self).__name__)
            self.__dict__["logger_"] = lg
        return lg

    def debug(self, msg, *args, **kw):
        self.logger.debug(msg, *args, **kw)

    def info(self, msg, *args, **kw):
        self.logger.info(msg, *args, **kw)

    def warning(self, msg, *args, **kw):
        self.logger.warning(msg, *args, **kw)

    def error(self, msg, *args, **kw):
        self.logger.error(msg, *args, **kw)

    def exception(self, msg="Exception", *args, **kw):
        self.logger.exception(msg, *args, **kw)
