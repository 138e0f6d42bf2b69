"""Adds the `hrnet_small_c19` entry of tests/golden/param_shapes.json (BASELINE.json configs[3]:
"higher_hrnet 19-class ... mean-teacher EMA"): the parameter shapes of the reference's own
models/higher_hrnet.py:get_pose_net(POSE_HIGHER_RESOLUTION_NET) with NUM_JOINTS = 19, built in the
container where /root/reference is mounted.  `yacs` is not installed here; the reference only needs
attribute-style nested config nodes, so a 10-line stand-in is injected before the import.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_param_shapes_hrnet.py
"""
import json
import math
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))


class CfgNode(dict):
    def __init__(self, *a, new_allowed=False, **k):
        super().__init__(*a, **k)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def main():
    warnings.simplefilter("ignore")
    yacs, cfgm = types.ModuleType("yacs"), types.ModuleType("yacs.config")
    cfgm.CfgNode = CfgNode
    yacs.config = cfgm
    sys.modules["yacs"], sys.modules["yacs.config"] = yacs, cfgm
    sys.path.insert(0, "/root/reference")
    from models.higher_hrnet import get_pose_net, POSE_HIGHER_RESOLUTION_NET as cfg
    cfg.NUM_JOINTS = 19
    model = get_pose_net(cfg)
    shapes = [list(p.shape) for p in model.parameters()]
    path = os.path.join(HERE, "param_shapes.json")
    with open(path) as f:
        d = json.load(f)
    d["hrnet_small_c19"] = {"how": "models.higher_hrnet.get_pose_net(POSE_HIGHER_RESOLUTION_NET) with NUM_JOINTS=19",
                            "n_params": sum(math.prod(s) for s in shapes), "n_tensors": len(shapes),
                            "n_buffers": len(list(model.buffers())), "shapes": shapes}
    with open(path, "w") as f:
        json.dump(d, f)
    print(d["hrnet_small_c19"]["n_params"], len(shapes))


if __name__ == "__main__":
    main()
