"""Where does a free-running step leave the oracle?  Per step and layer: sign decisions that differ and the
oracle's |BN output| there.  usage: python tools/debug_parity.py [config1|config2] [grid]"""
import sys
from pathlib import Path
import numpy as np
REPO = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(REPO / "deep-neural-net_b200"), str(REPO / "oracle"), str(REPO / "tests")]
import bnn_oracle as O
import torch
from test_config_parity_gpu import build_net

which = sys.argv[1] if len(sys.argv) > 1 else "config1"
grid = len(sys.argv) > 2
units, B, steps = ([784, 256, 256, 10], 256, 3) if which == "config1" else ([784, 4096, 4096, 4096, 10], 8192, 2)
rng = np.random.default_rng(0)
params = O.new_params(units, rng)
batches = []
for _ in range(steps):
    if grid:
        X = (rng.integers(-16, 17, (B, units[0])) / 16.0).astype(np.float32)
    else:
        X = rng.random((B, units[0]), dtype=np.float32) * 2 - 1
    Y = np.eye(units[-1], dtype=np.int8)[rng.integers(0, units[-1], B)]
    batches.append((X, Y))
nn = build_net(units, params, 0.01)
ref = O.copy_params(params)
n = len(units) - 1
for s, (X, Y) in enumerate(batches):
    rec = {}
    probe = O.copy_params(ref)
    O.train_step(probe, X.copy(), Y, 0.01, record=rec)
    a = X
    for i, layer in enumerate(nn._layers):
        out = layer.forward(a, isTrain=True)
        z = layer.cache["z"][:, :units[i + 1]].float().cpu().numpy()
        zr = rec["caches"][i]["z_pre"]
        zerr = np.abs(z - zr).max() / np.abs(zr).max()
        if i < n - 1:
            got = out.t.cpu().numpy()
            want = rec["caches"][i + 1]["input"]
            bad = got != want
            bn = rec["caches"][i]["z"]
            print(f"step {s} layer {i}: z max err {zerr:.2e}; {int(bad.sum())} flips of {bad.size}; "
                  f"|BN out| at flips {np.abs(bn[bad])[:8]}; min |BN out| {np.abs(bn).min():.2e}")
        else:
            print(f"step {s} layer {i}: z max err {zerr:.2e}; logits err "
                  f"{np.abs(out.logits.cpu().numpy() - rec['caches'][i]['z']).max():.2e}")
        a = out
    for layer in nn._layers:
        layer.cache.clear()
    cost = nn.train_step(X, Y, 0.01)
    cost_ref = O.train_step(ref, X.copy(), Y, 0.01)
    print(f"step {s}: loss {cost:.7f} oracle {cost_ref:.7f} rel {abs(cost - cost_ref) / abs(cost_ref):.2e}")
    for i in range(n):
        w, wr = nn.weights[i].get(), ref["W"][i]
        print(f"   W{i} rel {np.linalg.norm(w - wr) / np.linalg.norm(wr):.2e}, sign(W) differs at "
              f"{int(((w >= 0) != (wr >= 0)).sum())}, bitwise different {int((w != wr).sum())} of {w.size}")
