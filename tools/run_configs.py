"""Runs the large BASELINE configurations on one GPU and prints one JSON line each (development / evidence aid):
  C3  100k x 400 aa triangle (4 999 950 000 pairs), results kept on the device as uint16 (10 GB), size-independent
      checks: symmetry against row mode, bound by min length, oracle on sampled pairs.
  C5  3M x 250 aa: the medoid path's shape -- 100 seed rows against all 3M sequences (row mode).
usage: python tools/run_configs.py [c3] [c5] [--n-scale 1.0]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import famsa_b200
from famsa_b200 import seqio
from oracle import pyoracle

scale = 1.0
if "--n-scale" in sys.argv:
    scale = float(sys.argv[sys.argv.index("--n-scale") + 1])
eng = famsa_b200.Engine(0)

if "c3" in sys.argv:
    n = int(100000 * scale)
    t = time.time(); codes, offsets, lens = seqio.synth_family(n, 400, seed=2); tg = time.time() - t
    t = time.time(); eng.upload(codes, offsets, lens); tu = time.time() - t
    pairs = n * (n - 1) // 2
    d = torch.empty(pairs, dtype=torch.int16, device="cuda")
    eng.triangle_device(0, n, d.data_ptr(), 2)
    tot, main, p = eng.last_timing()
    rng = np.random.default_rng(0)
    refs = rng.integers(1, n, size=4)
    rows = eng.rows(refs, dtype=np.uint16)
    ok = True
    for r, ref in enumerate(refs):
        ref = int(ref)
        below = d[ref * (ref - 1) // 2: ref * (ref - 1) // 2 + ref].cpu().numpy().view(np.uint16)
        ok &= bool(np.array_equal(rows[r, :ref], below)) and rows[r, ref] == lens[ref]
        cols = rng.integers(0, ref, size=100)
        want = pyoracle.lcs_rows(codes, offsets, lens, [ref], cols)[0]
        ok &= bool(np.array_equal(below[cols], want))
    print(json.dumps({"config": f"C3 {n} x 400 aa triangle on 1 B200", "pairs": pairs, "kernel_ms": main, "total_ms": tot,
                      "pairs_per_s": pairs / (main / 1e3), "upload_s": tu, "gen_s": tg, "checks_ok": bool(ok),
                      "checksum": int(d.to(torch.int64).sum().item())}))
    del d
    torch.cuda.empty_cache()

if "c5" in sys.argv:
    n = int(3000000 * scale)
    t = time.time(); codes, offsets, lens = seqio.synth_family(n, 250, seed=3, n_subroots=300); tg = time.time() - t
    t = time.time(); eng.upload(codes, offsets, lens); tu = time.time() - t
    rng = np.random.default_rng(1)
    seeds = np.sort(rng.choice(n, size=100, replace=False)).astype(np.uint32)
    d_ref = torch.from_numpy(seeds.astype(np.int32)).cuda()
    d_out = torch.empty(100 * n, dtype=torch.int16, device="cuda")
    eng.rows_device(d_ref.data_ptr(), 100, 0, n, d_out.data_ptr(), 2)
    tot, main, p = eng.last_timing()
    out = d_out.view(100, n)
    ok = True
    for r in (0, 57, 99):
        cols = rng.integers(0, n, size=200)
        want = pyoracle.lcs_rows(codes, offsets, lens, [int(seeds[r])], cols)[0]
        got = out[r, torch.from_numpy(cols).cuda()].cpu().numpy().view(np.uint16)
        ok &= bool(np.array_equal(got, want))
    t = time.time(); a, dmin = eng.assign(seeds, 0); t_assign = time.time() - t
    tot_a, main_a, _ = eng.last_timing()
    chk = rng.integers(0, n, size=50)
    for j in chk:
        ds = np.array([pyoracle.transform(0, int(out[r, int(j)].item()) & 0xffff, int(lens[int(seeds[r])]), int(lens[int(j)]), False) for r in range(100)], dtype=np.float32)
        ok &= bool(a[j] == int(np.argmin(ds)) and dmin[j] == ds.min())
    print(json.dumps({"config": f"C5 medoid assignment (famsa_lcs_assign): 100 seeds x {n} x 250 aa, returns {n} assignments",
                      "device_ms": tot_a, "wall_s_incl_D2H": t_assign, "checks_ok": bool(ok)}))
    print(json.dumps({"config": f"C5 shape: 100 seed rows x {n} x 250 aa on 1 B200", "pairs": 100 * n, "kernel_ms": main,
                      "total_ms": tot, "pairs_per_s": 100 * n / (tot / 1e3), "upload_s": tu, "gen_s": tg, "checks_ok": bool(ok)}))
