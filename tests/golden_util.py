import json
import os

import numpy as np

import oracle_lib as ol

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.json")


def load():
    with open(PATH) as f:
        g = json.load(f)
    for c in g["cases"]:
        c["draws"] = np.array([float.fromhex(h) for h in c["draws_hex"]]).reshape(c["draws_shape"])
        c["settings"] = ol.Settings(**c["st"])
        if "lower" in c:   # box constraints, stored as hex strings ("inf"/"-inf" = open side)
            c["settings"]["lower_bounds"] = np.array([float.fromhex(h) for h in c["lower"]])
            c["settings"]["upper_bounds"] = np.array([float.fromhex(h) for h in c["upper"]])
    return g
