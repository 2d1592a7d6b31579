"""one row per BASELINE.json configuration on 1 GPU (device-resident conversion,
CUDA events inside m2s_convert, median of 7 after 3 warm-ups, L2 not flushed here — bench.py is the
contract measurement for config 2).  Prints markdown rows + a JSON blob."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context

PEAK = 6577.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def alg_bytes(scene, n, layout):
    maps = set()
    for p in scene.primitives:
        ids = [p.albedo_texture] if layout == _abi.LAYOUT_PACKED56 else [p.albedo_texture, p.normal_texture, p.metallic_roughness_texture]
        maps.update(i for i in ids if i >= 0)
    return n * _abi.STRIDES[layout] + scene.triangle_count * 144 + sum(min(4 * n, scene.textures[i].nbytes) for i in maps)


def run(ctx, name, scene, R, layout, flags, parity=True):
    ds = ctx.upload(scene)
    cap = 6 * R * R * max(1, len(scene.primitives)) if flags & _abi.FLAG_UNCAPPED else _abi.reference_capacity(R, len(scene.primitives))
    cap = min(cap, 60_000_000)
    out, ts = None, []
    for i in range(10):
        out = ctx.convert(ds, R, layout, flags=flags, capacity=cap, out=out.data if out else None)
        ts.append(out.device_ms)
    t = float(np.median(ts[3:]))
    n = out.written
    row = {"config": name, "R": R, "layout": "PACKED56" if layout == 1 else "REF96", "triangles": scene.triangle_count,
           "primitives": len(scene.primitives), "total": out.total, "written": n, "ms": t, "Mg_s": n / t / 1e3,
           "alg_MB": alg_bytes(scene, n, layout) / 1e6, "GB_s": alg_bytes(scene, n, layout) / t / 1e6}
    row["frac"] = row["GB_s"] / PEAK
    if parity:
        t0 = time.perf_counter()
        prep = oracle.Prepared(scene)
        cn, ctot, _ = prep.convert(R, layout, flags=flags, capacity=cap)
        row["oracle_total"] = ctot
        row["count_match"] = bool(ctot == out.total)
        row["cpu_s"] = time.perf_counter() - t0
    ds.free()
    print(f"| {name} | 1 | {out.total:,} ({n:,} stored) | {t*1e3:.1f} µs | {row['Mg_s']:.0f} | {row['alg_MB']:.1f} MB | {row['GB_s']:.0f} | {100*row['frac']:.1f} % | "
          f"{'count == oracle' if row.get('count_match') else ('count != oracle' if parity else 'n/a')} |", flush=True)
    return row


ctx = Context(0)
rows = []
P, Rf = _abi.LAYOUT_PACKED56, _abi.LAYOUT_REF96
U = _abi.FLAG_UNCAPPED
quad = synth.unit_quad()
rows.append(run(ctx, "1 unit quad R=64 (REF96)", quad, 64, Rf, 0))
helmet = synth.helmet_standin(2048)
rows.append(run(ctx, "2 helmet stand-in R=512 (PACKED56)", helmet, 512, P, U))
rows.append(run(ctx, "2 helmet stand-in R=512 (REF96)", helmet, 512, Rf, U))
sponza = synth.sponza_standin(1024)
rows.append(run(ctx, "3 sponza stand-in R=1024, reference cap 7M (REF96)", sponza, 1024, Rf, 0))
rows.append(run(ctx, "3 sponza stand-in R=1024, uncapped (PACKED56)", sponza, 1024, P, U))
s1m = synth.sphere_1m(2048)
rows.append(run(ctx, "4 1M-triangle sphere R=256 (PACKED56)", s1m, 256, P, U))
dh = synth.damaged_helmet_standin(2048)
for R in (64, 128, 256, 512, 1024, 2048):
    rows.append(run(ctx, f"5 damaged-helmet stand-in R={R} (PACKED56)", dh, R, P, U, parity=R <= 1024))
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "config_sweep.json"), "w"), indent=1)
