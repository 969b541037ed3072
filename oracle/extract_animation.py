#!/usr/bin/env python3
"""TEST/BUILD INFRASTRUCTURE — extract the reference's animation DATA for the N1 frame-setup module.

The hard-coded animation of the reference is data embedded in scene.cc: a table of 249
`animation_stop{start, duration, from, to, &variable}` rows (scene.cc:319-627) and the name ->
(mesh, bvh) table `scene::meshes` (scene.hh:49) built by load_scene(). This script parses the rows
out of the reference source where it is mounted and asks the oracle for the mesh table, and writes
both to scenes/_cache/animation.json (git-ignored, travels to the GPU box). Nothing is copied into
the repository; the player that consumes the data is this repo's own (csrc/frame_setup.cu).
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# variable order = enum ptgpu_anim_var in include/ptgpu.h
VARS = ["logo_visible", "armadillo_visible", "dragon_visible", "bunny_visible", "end_visible",
        "cam.position.x", "cam.position.y", "cam.position.z",
        "cam_orientation.x", "cam_orientation.y", "cam_orientation.z",
        "fov", "cam.focal_distance", "cam.aperture_radius",
        "teapot_pos.x", "teapot_pos.y", "teapot_pos.z", "teapot_ori.x", "teapot_ori.y", "teapot_ori.z",
        "armadillo_pos.x", "armadillo_pos.y", "armadillo_pos.z", "armadillo_ori.x", "armadillo_ori.y", "armadillo_ori.z",
        "dragon_pos.x", "dragon_pos.y", "dragon_pos.z", "dragon_ori.x", "dragon_ori.y", "dragon_ori.z",
        "bunny_pos.x", "bunny_pos.y", "bunny_pos.z", "bunny_ori.x", "bunny_ori.y", "bunny_ori.z",
        "end_pos.x", "end_pos.y", "end_pos.z", "end_ori.x", "end_ori.y", "end_ori.z"]
MESHES = ["logo", "buddha", "teapot", "armadillo", "dragon", "bunny", "end"]


def parse_keys(scene_cc):
    text = open(scene_cc).read()
    a = text.index("const animation_stop anim[] = {")
    b = text.index("};", a)
    body = text[a:b]
    consts = {}
    for name in ("camera_start_pos", "camera_start_ori"):
        m = re.search(r"const float3 %s = float3\{([^}]*)\}" % name, text)
        vals = [float(x) for x in m.group(1).split(",")]
        for c, v in zip("xyz", vals):
            consts["%s.%s" % (name, c)] = v
    keys = []
    for m in re.finditer(r"\{([^{}]*?),\s*&([A-Za-z_.]+)\s*\}", body):
        fields = [f.strip() for f in m.group(1).split(",")]
        if len(fields) != 4:
            raise ValueError("unexpected row: %r" % m.group(0))
        nums = []
        for f in fields:
            f = f.rstrip("f") if re.fullmatch(r"-?[0-9.]+f", f) else f
            nums.append(consts[f] if f in consts else float(f))
        keys.append(nums + [VARS.index(m.group(2))])
    return keys, consts


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = os.path.join(ROOT, "scenes", "_cache", "animation.json")
    src = os.path.join(ref, "scene.cc")
    if not os.path.exists(src):
        print("extract_animation: %s not present; keeping %s" % (src, out))
        return 0 if os.path.exists(out) else 1
    keys, consts = parse_keys(src)
    from oracle import refbind
    o = refbind.get("fast")
    o.load_scene()
    meshes = {}
    for name in MESHES:
        m = o.find_mesh(name)  # vertex_count, triangle_count, index_offset, base_vertex_offset, node_count, node_offset
        meshes[name] = {"mesh": m[:4], "blas": m[4:6]}
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump({"vars": VARS, "keys": keys, "meshes": meshes, "n_keys": len(keys),
                   "source": "reference scene.cc:319-627 (animation_stop table) and scene::meshes after load_scene()"}, f)
    print("extract_animation: %d keys, %d meshes -> %s" % (len(keys), len(meshes), out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
