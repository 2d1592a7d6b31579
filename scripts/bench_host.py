"""Host-side rows next to the hot path, timed against the REFERENCE's own code compiled here
(oracle/_ref: SceneManager::parseGltfFile + tinygltf + stb_image; parsers::savePlyVector):

  loader   .glb (70 074 triangles, three 2048^2 maps as PNG / JPEG) -> host scene
  writer   643 k REF96 records -> .ply (formats 0, 1, 2)

usage: python scripts/bench_host.py [--tex 2048] [--reps 3]      (CPU only; needs /root/reference for the reference arm)
"""
import argparse, io, json, os, struct, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.gltf import load_glb
from mesh2splat_b200.api import ply_write


def write_glb(path, scene, image_blobs):
    """Non-indexed single-primitive .glb of a Scene with embedded images [(bytes, mime)] for albedo, normal, MR."""
    t = scene.triangles.reshape(-1, 3, 12)
    pos, nrm, tan, uv = (np.ascontiguousarray(t[:, :, a:b].reshape(-1, b - a)) for a, b in ((0, 3), (3, 6), (6, 10), (10, 12)))
    blobs, views, accessors = [], [], []

    def add_view(b):
        off = sum(len(x) for x in blobs); pad = (-len(b)) % 4
        blobs.append(b + b"\x00" * pad); views.append({"buffer": 0, "byteOffset": off, "byteLength": len(b)}); return len(views) - 1

    def add_acc(arr, typ):
        accessors.append({"bufferView": add_view(arr.tobytes()), "componentType": 5126, "count": len(arr), "type": typ}); return len(accessors) - 1

    attrs = {"POSITION": add_acc(pos, "VEC3"), "NORMAL": add_acc(nrm, "VEC3"), "TANGENT": add_acc(tan, "VEC4"), "TEXCOORD_0": add_acc(uv, "VEC2")}
    imgs = [{"bufferView": add_view(b), "mimeType": m} for b, m in image_blobs]
    gltf = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
            "meshes": [{"name": "helmet", "primitives": [{"attributes": attrs, "material": 0}]}],
            "materials": [{"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}, "metallicRoughnessTexture": {"index": 2}},
                           "normalTexture": {"index": 1}}],
            "textures": [{"source": 0}, {"source": 1}, {"source": 2}], "images": imgs, "bufferViews": views, "accessors": accessors}
    binblob = b"".join(blobs); gltf["buffers"] = [{"byteLength": len(binblob)}]
    js = json.dumps(gltf).encode(); js += b" " * ((-len(js)) % 4)
    with open(path, "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, 12 + 8 + len(js) + 8 + len(binblob)))
        f.write(struct.pack("<I4s", len(js), b"JSON")); f.write(js)
        f.write(struct.pack("<I4s", len(binblob), b"BIN\x00")); f.write(binblob)


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--tex", type=int, default=2048); ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from PIL import Image
    scene = synth.helmet_standin(a.tex)
    have_ref = oracle.ref_loader_lib() is not None
    d = tempfile.mkdtemp()
    rows = []
    for label, fmts in (("3 x PNG", ("PNG", "PNG", "PNG")), ("PNG albedo + 2 x JPEG q90 4:2:0", ("PNG", "JPEG", "JPEG")), ("3 x progressive JPEG", ("PJPEG",) * 3)):
        blobs = []
        for tex, fmt in zip(scene.textures, fmts):
            b = io.BytesIO()
            if fmt == "PNG": Image.fromarray(tex).save(b, "PNG", compress_level=6); blobs.append((b.getvalue(), "image/png"))
            else: Image.fromarray(tex[..., :3]).save(b, "JPEG", quality=90, subsampling=2, progressive=(fmt == "PJPEG")); blobs.append((b.getvalue(), "image/jpeg"))
        path = os.path.join(d, "helmet.glb"); write_glb(path, scene, blobs)
        size = os.path.getsize(path) / 1e6
        t_ours = best(lambda: load_glb(path), a.reps)
        t_ref = best(lambda: oracle.ref_load_glb(path), a.reps) if have_ref else None
        s = load_glb(path)
        if have_ref:
            ok, meshes = oracle.ref_load_glb(path)
            same = np.array_equal(np.vstack([m["faces"] for m in meshes]).view(np.uint32), s.triangles.view(np.uint32)) and all(
                np.array_equal(meshes[0]["textures"][k], s.textures[i]) for k, i in ((0, s.primitives[0].albedo_texture), (1, s.primitives[0].normal_texture), (2, s.primitives[0].metallic_roughness_texture)))
        else:
            same = None
        rows.append(("loader: " + label, f"{size:.1f} MB file", t_ours, t_ref, same))
    # writer
    rng = np.random.default_rng(1)
    n = 643438
    rec = rng.random((n, 24)).astype(np.float32); rec[:, 8:11] *= 0.01
    recs = rec.view(_abi.record_dtype(_abi.LAYOUT_REF96)).reshape(-1)
    mult = float(np.float32(0.65) / np.float32(512))
    for fmt in (0, 1, 2):
        p1, p2 = os.path.join(d, "ours.ply"), os.path.join(d, "ref.ply")
        t_ours = best(lambda: ply_write(p1, recs, fmt, mult), a.reps)
        t_ref = best(lambda: oracle.ref_save_ply(p2, rec, fmt, mult), a.reps) if oracle.ref_ply_lib() is not None else None
        same = (open(p1, "rb").read() == open(p2, "rb").read()) if t_ref is not None else None
        rows.append((f"writer: format {fmt}", f"{os.path.getsize(p1) / 1e6:.0f} MB file, {n} gaussians", t_ours, t_ref, same))
    print("| row | input | ours (s) | reference's own code (s) | speed-up | identical output |\n|---|---|---|---|---|---|")
    for name, inp, to, tr, same in rows:
        print(f"| {name} | {inp} | {to:.3f} | {'n/a' if tr is None else f'{tr:.3f}'} | {'n/a' if tr is None else f'{tr / to:.2f}x'} | {same} |")


if __name__ == "__main__":
    main()
