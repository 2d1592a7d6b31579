"""Mutation fuzz of m2s_glb_load: corrupt golden .glb fixtures (bit flips, truncation, length-field edits) and load
them in a child process; any crash (signal) is reported with the seed that reproduces it."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from mesh2splat_b200.gltf import load_glb
g = np.load(%r)
names = [str(n) for n in g["names"]]
import os, tempfile
seed0, count = int(sys.argv[1]), int(sys.argv[2])
d = tempfile.mkdtemp()
for seed in range(seed0, seed0 + count):
    rng = np.random.default_rng(seed)
    blob = bytearray(g[names[seed %% len(names)] + "/glb"].tobytes())
    kind = seed %% 5
    if kind == 0:
        for _ in range(int(rng.integers(1, 8))):
            blob[int(rng.integers(0, len(blob)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:
        blob = blob[: int(rng.integers(0, len(blob)))]
    elif kind == 2:
        i = int(rng.integers(0, max(1, len(blob) - 4)))
        blob[i:i + 4] = int(rng.integers(0, 2**32)).to_bytes(4, "little")
    elif kind == 4:
        # JSON-level: replace one numeric literal of the JSON chunk by a hostile value (negative, huge, fractional);
        # keeps the chunk length by padding/truncating with spaces when possible (ADVICE r1: negative byteOffset)
        import re, struct
        jlen = struct.unpack_from("<I", blob, 12)[0]
        js = bytes(blob[20:20 + jlen])
        nums = list(re.finditer(rb"(?<=[:\[,\s])-?\d+(\.\d+)?", js))
        if nums:
            m = nums[int(rng.integers(0, len(nums)))]
            bad = [b"-1", b"-3", b"-1000000", b"2147483648", b"1e300", b"0.5", b"-0.5", b"1099511627776"][int(rng.integers(0, 8))]
            js2 = js[:m.start()] + bad + js[m.end():]
            js2 += b" " * ((-len(js2)) %% 4)
            rest = bytes(blob[20 + jlen:])
            blob = bytearray(struct.pack("<4sII", b"glTF", 2, 20 + len(js2) + len(rest)) + struct.pack("<I4s", len(js2), b"JSON") + js2 + rest)
    else:
        i = int(rng.integers(0, len(blob))); j = int(rng.integers(0, len(blob)))
        blob[i], blob[j] = blob[j], blob[i]
        blob[int(rng.integers(0, len(blob)))] = int(rng.integers(0, 256))
    p = os.path.join(d, "f.glb")
    open(p, "wb").write(bytes(blob))
    print(seed, flush=True)
    try:
        load_glb(p)
    except (ValueError, OSError):
        pass
print("done", flush=True)
''' % (ROOT, os.path.join(ROOT, "tests", "golden", "ref_loader_vectors.npz"))

def run(seed0, count):
    r = subprocess.run([sys.executable, "-c", CHILD, str(seed0), str(count)], capture_output=True, text=True, timeout=600)
    lines = r.stdout.strip().splitlines()
    if r.returncode != 0 or not lines or lines[-1] != "done":
        last = lines[-1] if lines else "?"
        return False, f"crash/abort rc={r.returncode} at seed {last}: {r.stderr.strip()[-300:]}"
    return True, ""

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for s0 in range(seed, seed + n, 500):
        ok, msg = run(s0, min(500, seed + n - s0))
        if not ok:
            bad += 1; print(msg)
    print("batches with crashes:", bad)
