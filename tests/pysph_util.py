"""Helpers for the pysplishsplash tests: a small scene in the reference's JSON schema plus a unit-box OBJ."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

UNIT_BOX_OBJ = """# unit box centred at the origin
v -0.5 -0.5 -0.5
v 0.5 -0.5 -0.5
v 0.5 0.5 -0.5
v -0.5 0.5 -0.5
v -0.5 -0.5 0.5
v 0.5 -0.5 0.5
v 0.5 0.5 0.5
v -0.5 0.5 0.5
f 1 2 3 4
f 5 8 7 6
f 1 5 6 2
f 2 6 7 3
f 3 7 8 4
f 4 8 5 1
"""


def import_sph():
    p = os.path.join(ROOT, "difffr_b200")
    if p not in sys.path:
        sys.path.insert(0, p)
    import pysplishsplash as sph

    return sph


def write_scene(tmp_path, target_time=0.02, manager=False, extra_cfg=None):
    """Tank 1 x 0.8 x 0.6 with a fluid block and one dynamic box (boundary model 1), keys as in the reference's
    experiments/rigid_body_trajectory_optimization/scene/*.json."""
    os.makedirs(tmp_path / "models", exist_ok=True)
    os.makedirs(tmp_path / "scene", exist_ok=True)
    (tmp_path / "models" / "UnitBox.obj").write_text(UNIT_BOX_OBJ)
    cfg = {
        "particleRadius": 0.025, "simulationMethod": 5, "gravitation": [0, -9.81, 0], "cflMethod": 1, "cflFactor": 0.5,
        "cflMaxTimeStepSize": 0.005, "maxIterations": 100, "maxError": 0.05, "maxIterationsV": 100, "maxErrorV": 0.1,
        "enableDivergenceSolver": True, "boundaryHandlingMethod": 0, "targetTime": target_time, "uniformAccelerateRBTime": 0.0,
        "useRigidGradientManager": bool(manager),
    }
    cfg.update(extra_cfg or {})
    scene = {
        "Configuration": cfg,
        "Materials": [{"id": "Fluid", "surfaceTension": 0.2, "surfaceTensionBoundary": 0.0, "surfaceTensionMethod": 2}],
        "RigidBodies": [
            {"id": 1, "geometryFile": "../models/UnitBox.obj", "translation": [0, 0.4, 0], "rotationAxis": [1, 0, 0], "rotationAngle": 0,
             "scale": [1.0, 0.8, 0.6], "isDynamic": False, "isWall": True},
            {"id": 2, "geometryFile": "../models/UnitBox.obj", "isDynamic": 1, "density": 500, "translation": [-0.15, 0.42, 0.0],
             "rotationAxis": [0, 0, 1], "rotationAngle": 0.2, "scale": [0.2, 0.2, 0.2], "initVelocity": [0.5, -0.2, 0.1],
             "initAngularVelocity": [0.3, 1.0, -0.4], "targetX": [0.2, 0.3, 0.0], "targetAngleInDegree": [0, 30, 0]},
        ],
        "FluidBlocks": [{"denseMode": 0, "start": [-0.5, 0.0, -0.3], "end": [0.1, 0.35, 0.3], "translation": [0, 0, 0], "scale": [1, 1, 1]}],
    }
    path = tmp_path / "scene" / "mini.json"
    path.write_text(json.dumps(scene, indent=1))
    return str(path)
