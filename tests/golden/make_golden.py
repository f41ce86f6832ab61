"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference/ModularER_2D) under tests/golden/ref_shim.py. Build container only.

    python tests/golden/make_golden.py

Outputs (committed):
  terrain.npz                 terrain_x/terrain_y for env.seed(4), rough and flat           (pin P1)
  morphology_<enc>.npz        body/joint/controller tables recorded from the reference's own
                              Modular2D.reset()/create_robot for seeded random individuals,
                              fresh and after mutations, enc in {direct, lsystem, ce}        (pin P2)
  control_pin.json            motorSpeed written by the reference's step() on frozen bodies   (a9/a10)
The recording fake Box2D only rounds what pybox2d stores as float32 (see ref_shim.py).
"""
import json
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import ref_shim  # noqa: E402

r2d = ref_shim.install()

N_FRESH = {"direct": 300, "lsystem": 300, "ce": 200}
N_MUTATED = 40          # first N_MUTATED individuals are additionally recorded after 3 mutation rounds
MUT_ARGS = (0.3, 0.3, 0.2)


def record(env, individual):
    tree = individual.genome.create(individual.tree_depth)
    env.seed(4)
    env.reset(tree=tree, module_list=individual.genome.moduleList)
    w = env.world
    assert len(w.static_bodies) == 199
    bodies = w.dynamic_bodies
    slot = {id(b): i for i, b in enumerate(bodies)}
    rec = dict(shape=[], hx=[], hy=[], x0=[], y0=[], a0=[], joint_parent=[], anchor_a=[], anchor_b=[],
               lower=[], upper=[], max_torque=[], ctrl=[], node_index=[], type_ref=[])
    for b in bodies:
        if b.shape_kind == "circle":
            rec["shape"].append(1); rec["hx"].append(np.float32(b.radius)); rec["hy"].append(np.float32(0))
        else:
            rec["shape"].append(0); rec["hx"].append(np.float32(b.box[0])); rec["hy"].append(np.float32(b.box[1]))
        assert b.fixture_kw["density"] == 1 and b.fixture_kw["friction"] == 0.1
        assert b.fixture_kw["categoryBits"] == 0x20 and b.fixture_kw["maskBits"] == 0x1
        rec["x0"].append(np.float32(b.position.x)); rec["y0"].append(np.float32(b.position.y))
        rec["a0"].append(np.float32(b.angle))
    for j in w.joints:
        assert slot[id(j.bodyB)] == len(rec["joint_parent"]) + 1
        rec["joint_parent"].append(slot[id(j.bodyA)])
        rec["anchor_a"].append([np.float32(v) for v in j.kw["localAnchorA"]])
        rec["anchor_b"].append([np.float32(v) for v in j.kw["localAnchorB"]])
        rec["lower"].append(np.float32(j.kw["lowerAngle"])); rec["upper"].append(np.float32(j.kw["upperAngle"]))
        rec["max_torque"].append(np.float32(j.kw["maxMotorTorque"]))
        assert j.kw["enableMotor"] is True and j.kw["enableLimit"] is True and "referenceAngle" not in j.kw
    for n in env.tree_morphology.nodes:
        if n.expressed and n.component is not None:
            c = n.controller
            rec["ctrl"].append([c.amplitude, c.phase, c.frequency, c.offset, c.i_state])
            rec["node_index"].append(n.index); rec["type_ref"].append(n.type)
    assert len(rec["ctrl"]) == len(bodies)
    return rec


def make_morphology(enc):
    env = ref_shim.reference_env()
    recs = []
    random.seed(1000 + len(enc)); np.random.seed(1000 + len(enc))
    seeds = []
    for i in range(N_FRESH[enc]):
        seed = 7919 * (i + 1) + len(enc)
        seeds.append(seed)
        random.seed(seed)
        ind = r2d.Individual.random(encoding=enc)
        recs.append(record(env, ind))
        if i < N_MUTATED:
            for _ in range(3):
                ind.genome.mutate(*MUT_ARGS)
            recs.append(record(env, ind))
    nb = np.array([len(r["shape"]) for r in recs], np.int32)
    out = dict(seeds=np.array(seeds, np.int64), n_mutated=np.int64(N_MUTATED), mut_args=np.array(MUT_ARGS),
               body_off=np.concatenate([[0], np.cumsum(nb)]).astype(np.int32))
    for k, dt in (("shape", np.uint8), ("hx", np.float32), ("hy", np.float32), ("x0", np.float32), ("y0", np.float32),
                  ("a0", np.float32), ("joint_parent", np.int16), ("lower", np.float32), ("upper", np.float32),
                  ("max_torque", np.float32), ("node_index", np.int32), ("type_ref", np.int16)):
        out[k] = np.array([v for r in recs for v in r[k]], dt)
    out["anchor_a"] = np.array([v for r in recs for v in r["anchor_a"]], np.float32).reshape(-1, 2)
    out["anchor_b"] = np.array([v for r in recs for v in r["anchor_b"]], np.float32).reshape(-1, 2)
    out["ctrl"] = np.array([v for r in recs for v in r["ctrl"]], np.float64).reshape(-1, 5)
    np.savez_compressed(os.path.join(HERE, "morphology_%s.npz" % enc), **out)
    print(enc, "creatures", len(recs), "bodies", int(nb.sum()), "mean", nb.mean(), "max", nb.max())


def make_terrain():
    out = {}
    for name, flat in (("rough", False), ("flat", True)):
        env = ref_shim.reference_env(flat=flat)
        env.seed(4)
        env.reset(tree=None)
        out[name + "_x"] = np.array(env.terrain_x, np.float64)
        out[name + "_y"] = np.array(env.terrain_y, np.float64)
        edges = [b.vertices for b in env.world.static_bodies]
        out[name + "_edges"] = np.array(edges, np.float64)          # creation order = ascending x
        assert all(b.fixture_kw["friction"] == 2.5 and b.fixture_kw["categoryBits"] == 1 for b in env.world.static_bodies)
    np.savez_compressed(os.path.join(HERE, "terrain.npz"), **out)
    ref_shim.reference_env(flat=False)   # restore module constant


def make_control_pin():
    """Reference step() on frozen bodies: pins controller.update + P-control (a9, a10) and the WOD."""
    env = ref_shim.reference_env()
    random.seed(3)
    from Encodings import direct_encoding as de
    ind = r2d.Individual()
    ind.genome = de.DirectEncoding(r2d.get_module_list())
    ind.tree_depth = 8
    tree = ind.genome.create(8)
    env.seed(4)
    env.reset(tree=tree, module_list=ind.genome.moduleList)
    ticks = []
    for _ in range(5):
        obs, reward, done, info = env.step(None)
        ticks.append(dict(motor_speed=[float(np.float32(j.motorSpeed)) for j in env.world.joints],
                          motor_speed_f64=[float(j.motorSpeed) for j in env.world.joints],
                          wod=env.wod.position, reward=float(reward), done=bool(done)))
    pin = dict(seed=3, joint_angles=[j.angle for j in env.world.joints], ticks=ticks,
               step_args=list(env.world.last_step_args))
    with open(os.path.join(HERE, "control_pin.json"), "w") as f:
        json.dump(pin, f, indent=1)


if __name__ == "__main__":
    make_terrain()
    make_control_pin()
    for enc in ("direct", "lsystem", "ce"):
        make_morphology(enc)
