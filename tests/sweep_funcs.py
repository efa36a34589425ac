"""Module-level simulation functions for the replica-scheduler tests (workers import them by name)."""
import os
import time


def toy(params):
    if params["w"] == 3.0 and params["h"] == 20:
        raise ValueError("bad corner")
    time.sleep(0.05 * float(params["w"]))
    return {"metric": float(params["w"]) * 10 + float(params["h"]), "env_device": int(os.environ["PRISMO_B200_DEVICE"]),
            "pid": os.getpid()}


def configured_device(params):
    from prismo_b200 import session

    return {"dev": session.configure()["device"]}
