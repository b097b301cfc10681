"""Per-run log files (mirrors mcmc_logging.py of the reference: one timestamped file per logger)."""
import logging
import os
from datetime import datetime


def start_logging(log_dir="Logs", name="Log", verbose=False):
    os.makedirs(log_dir, exist_ok=True)
    stamp = datetime.now().strftime("%Y-%m-%d_%H-%M-%S-%f")
    logger = logging.getLogger(f"{name}{stamp}")
    logger.setLevel(logging.DEBUG if verbose else logging.INFO)
    handler = logging.FileHandler(os.path.join(log_dir, f"{name}{stamp}.log"))
    handler.setFormatter(logging.Formatter("%(asctime)s [%(levelname)s] %(message)s"))
    logger.addHandler(handler)
    return logger, handler


def stop_logging(logger, handler, err=0):
    if err:
        logger.error(f"Termination with error code {err}")
    logger.removeHandler(handler)
    handler.close()
    logging.shutdown()
