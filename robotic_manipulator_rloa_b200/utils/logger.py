"""
Host-side logging, behaviour-compatible with the reference
(/root/reference/robotic_manipulator_rloa/utils/logger.py:12-27, 62-108): a root dictConfig with an
ANSI-coloured stdout handler plus a 50 MB x 10 rotating ``training_logs.log`` in the CWD at INFO.
Never called per step on the device path.
"""
import logging
from datetime import datetime, timezone
from logging.config import dictConfig
from logging.handlers import RotatingFileHandler

_COLOURS = {logging.DEBUG: '\033[32;20m', logging.INFO: '\033[38;20m', logging.WARNING: '\033[33;20m',
            logging.ERROR: '\033[31;20m', logging.CRITICAL: '\033[31;1m'}
_RESET = '\033[0m'


class CustomFormatter(logging.Formatter):
    """``[LEVEL   ] - <local ISO time> - message`` coloured by level."""

    def __init__(self, dateformat: str = None):
        super().__init__()
        self.dateformat = dateformat

    def format(self, record: logging.LogRecord) -> str:
        stamp = datetime.now().astimezone().strftime('%Y-%m-%dT%H:%M:%S.%f%z')
        colour = _COLOURS.get(record.levelno, '')
        fmt = f'{colour}[%(levelname)-8s] - {stamp} - %(message)s{_RESET}'
        return logging.Formatter(fmt, datefmt=self.dateformat).format(record)


def get_global_logger() -> logging.Logger:
    return logging.getLogger(__name__)


class Logger:
    @staticmethod
    def generate_logging_config_dict() -> dict:
        return {
            'version': 1,
            'disable_existing_loggers': False,
            'formatters': {'custom_formatter': {'()': CustomFormatter, 'dateformat': '%Y-%m-%dT%H:%M:%S.%06d%z'}},
            'handlers': {'debug_console_handler': {'level': 'NOTSET', 'formatter': 'custom_formatter',
                                                   'class': 'logging.StreamHandler', 'stream': 'ext://sys.stdout'}},
            'loggers': {'': {'handlers': ['debug_console_handler'], 'level': 'NOTSET'}},
        }

    @staticmethod
    def set_logger_setup() -> None:
        dictConfig(Logger.generate_logging_config_dict())
        handler = RotatingFileHandler(filename='training_logs.log', mode='a', maxBytes=50000000, backupCount=10,
                                      encoding='utf-8')
        stamp = datetime.now(timezone.utc).strftime('%Y-%m-%dT%H:%M:%S.%fZ')
        handler.setFormatter(logging.Formatter(f'"%(levelname)s"|"{stamp}"|%(message)s'))
        handler.setLevel(logging.INFO)
        logger = get_global_logger()
        logger.addHandler(handler)
        logger.setLevel(20)
