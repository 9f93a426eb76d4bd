"""GPU-box diagnostic: one sequence, device step vs the oracle run on the device's pose, per-seed differences printed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from android_svo_b200 import capi, synth
from oracle.pyoracle import Oracle, Cam, OracleSeq
import scenes
from test_pipeline import make_sequence, seed_matrix

oracle = Oracle()
ctx = capi.Context(0)
cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE, 5)
cam_o, cam_g = scenes.cam_of(cfg, Cam), scenes.cam_of(cfg, capi.Camera)
args = (cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
N, S = cfg["n_features"], cfg["n_seeds"]
pin = OracleSeq(oracle, cam_o, *args)
pin.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"]); pin.set_last(imgs[0])
trk = capi.Tracker(ctx, cam_g, 1, *args)
trk.set_keyframe(imgs[0][None], poses[0][None], [0, N], kf["kf_px"], kf["kf_level"], kf["pt_world"], [0, S], kf["seed_px"], kf["seed_level"])
trk.set_last(imgs[0][None])
for k in range(1, 5):
    sg0 = seed_matrix(trk.seeds()).copy()
    st = trk.step(imgs[k][None], poses[k - 1][None], last_px[k - 1])
    pin.set_pose_override(st[0]["T_cur_w"])
    pin.step(imgs[k], poses[k - 1], last_px[k - 1])
    og, op = trk.seed_obs(), pin.seed_obs()
    sg, sp = seed_matrix(trk.seeds()), pin.seeds()
    bad = np.nonzero((sg.view(np.uint32) != sp.view(np.uint32)).any(axis=1))[0]
    print("frame %d: %d of %d seeds differ in bits; status diff %d, n_evals diff %d, z diff %d, px diff %d" % (
        k, len(bad), S, (og["status"] != op["status"]).sum(), (og["n_evals"] != op["n_evals"]).sum(), (og["z"] != op["z"]).sum(),
        (og["px_cur"] != op["px_cur"]).any(axis=1).sum()))
    for i in bad[:6]:
        print("  seed", i, "level", kf["seed_level"][i], "px", kf["seed_px"][i], "\n    before", sg0[i], "\n    gpu   ", sg[i], "\n    oracle", sp[i],
              "\n    obs gpu", og[i], "\n    obs ora", op[i])
trk.close(); pin.close(); ctx.close()
