#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — stage the scene assets next to the oracle.

load_scene() (reference scene.cc:139-182) reads 18 OBJ files through relative paths
`data/<name>.obj`. The mount holds 15 of them; `terrain.obj`, `pine_tree.obj` and `bunny.obj`
are absent (`/root/reference/.MISSING_LARGE_BLOBS`), while their `.mtl` files are present.
This script
  * copies the present `data/*.obj|*.mtl` (assets, not source code) into `--out`, and
  * writes deterministic stand-ins for the three missing OBJs (fixed formulas, no RNG), using the
    material names of the real `.mtl` files so that materials bind (mesh.cc:199-210).
Everything lands under oracle/_ref/data (git-ignored; travels to the GPU box). Every number this
repo reports is therefore on a STAND-IN scene for terrain / pine trees / bunny; DESIGN.md says so.
"""
import argparse
import hashlib
import math
import os
import shutil
import sys

N = 192          # terrain quads per side
EXTENT = 100.0   # terrain spans [-EXTENT, EXTENT] in x and z
WATER_Y = 2.0


def height(x, z):
    # Spans roughly -10..30 so that every branch of the placement logic in scene.cc:141-152 and
    # :230-242 (tropical < 10 <= deciduous < 20 <= pine; gradients up to 28) is exercised.
    return (10.0 + 12.0 * math.sin(0.045 * x) * math.cos(0.05 * z)
            + 8.0 * math.sin(0.021 * (x + z) + 1.0)
            + 3.0 * math.sin(0.2 * x) * math.sin(0.17 * z))


def normal(x, z):
    dhdx = (12.0 * 0.045 * math.cos(0.045 * x) * math.cos(0.05 * z)
            + 8.0 * 0.021 * math.cos(0.021 * (x + z) + 1.0)
            + 3.0 * 0.2 * math.cos(0.2 * x) * math.sin(0.17 * z))
    dhdz = (-12.0 * 0.05 * math.sin(0.045 * x) * math.sin(0.05 * z)
            + 8.0 * 0.021 * math.cos(0.021 * (x + z) + 1.0)
            + 3.0 * 0.17 * math.sin(0.2 * x) * math.cos(0.17 * z))
    nx, ny, nz = -dhdx, 1.0, -dhdz
    l = math.sqrt(nx * nx + ny * ny + nz * nz)
    return nx / l, ny / l, nz / l


def write_terrain(path):
    out = ["# stand-in terrain (oracle/gen_assets.py), heightfield %dx%d" % (N, N),
           "mtllib terrain.mtl", "o terrain"]
    for j in range(N + 1):
        for i in range(N + 1):
            x = -EXTENT + 2.0 * EXTENT * i / N
            z = -EXTENT + 2.0 * EXTENT * j / N
            out.append("v %.6f %.6f %.6f" % (x, height(x, z), z))
    # water quad vertices
    wbase = (N + 1) * (N + 1)
    for (x, z) in ((-EXTENT, -EXTENT), (EXTENT, -EXTENT), (EXTENT, EXTENT), (-EXTENT, EXTENT)):
        out.append("v %.6f %.6f %.6f" % (x, WATER_Y, z))
    for j in range(N + 1):
        for i in range(N + 1):
            x = -EXTENT + 2.0 * EXTENT * i / N
            z = -EXTENT + 2.0 * EXTENT * j / N
            out.append("vn %.6f %.6f %.6f" % normal(x, z))
    out.append("vn 0.000000 1.000000 0.000000")
    nup = wbase + 1
    out.append("usemtl Material.003")  # land (terrain.mtl: Pr 0.5, no Tf)
    for j in range(N):
        for i in range(N):
            a = j * (N + 1) + i + 1
            b = a + 1
            c = a + (N + 1)
            d = c + 1
            # counter-clockwise seen from +y, so the upper side is the front face
            out.append("f %d//%d %d//%d %d//%d" % (a, a, c, c, b, b))
            out.append("f %d//%d %d//%d %d//%d" % (b, b, c, c, d, d))
    out.append("usemtl Material.001")  # water (terrain.mtl: Tf 1 -> material.z != 0, scene.cc:118,159)
    w = [wbase + 1, wbase + 2, wbase + 3, wbase + 4]
    out.append("f %d//%d %d//%d %d//%d" % (w[0], nup, w[3], nup, w[1], nup))
    out.append("f %d//%d %d//%d %d//%d" % (w[1], nup, w[3], nup, w[2], nup))
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def rewrite_obj(src, dst, mtllib, material_map):
    """Copy an OBJ, replacing its mtllib and renaming usemtl names (geometry untouched)."""
    with open(src) as f, open(dst, "w") as g:
        for line in f:
            if line.startswith("mtllib"):
                g.write("mtllib %s\n" % mtllib)
            elif line.startswith("usemtl"):
                name = line.split()[1]
                g.write("usemtl %s\n" % material_map.get(name, name))
            else:
                g.write(line)


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "data"))
    args = ap.parse_args()
    src = os.path.join(args.ref, "data")
    if not os.path.isdir(src):
        print("gen_assets: %s not present; keeping existing %s" % (src, args.out))
        return 0 if os.path.isdir(args.out) else 1
    os.makedirs(args.out, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if name.endswith((".obj", ".mtl")):
            dst = os.path.join(args.out, name)
            if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(os.path.join(src, name)):
                shutil.copyfile(os.path.join(src, name), dst)
    write_terrain(os.path.join(args.out, "terrain.obj"))
    # pine tree: willow geometry bound to the pine materials (trunk .015 -> .011, foliage .014 -> .010)
    rewrite_obj(os.path.join(src, "willow_tree.obj"), os.path.join(args.out, "pine_tree.obj"),
                "pine_tree.mtl", {"Material.015": "Material.011", "Material.014": "Material.010"})
    # bunny: teapot geometry bound to the bunny material
    rewrite_obj(os.path.join(src, "teapot.obj"), os.path.join(args.out, "bunny.obj"),
                "bunny.mtl", {"Material.005": "Material.024"})
    with open(os.path.join(args.out, "STANDINS.sha256"), "w") as f:
        for name in ("terrain.obj", "pine_tree.obj", "bunny.obj"):
            f.write("%s  %s\n" % (sha256(os.path.join(args.out, name)), name))
    return 0


if __name__ == "__main__":
    sys.exit(main())
