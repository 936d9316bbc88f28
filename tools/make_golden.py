"""Generate the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the GPU box never sees the reference).

  atlas_models.npy   [29,512] u32 — the 29 block models of the 13 .vox files in the load order of
                     src/game.zig:101-113, converted by OUR .vox reader + atlas builder
                     (texel index x + 8*y + 64*z after the y/z swap of voxel.zig:106-108).
  atlas_counts.json  filled sub-voxels per model (cross-checked against SURVEY App. B.3).
  chicken_32.npy     [32768] u32 — assets/chicken.vox (the entity model of src/game.zig:114) as 32^3 texels, x + 32*(y + 32*z),
                     converted by voxel.load_model (same y/z swap and palette rule as the block models).
  glslref_*.npz      outputs of the REFERENCE'S OWN shader text (oracle/_ref/libglslref.so = assets/shaders/*.glsl compiled for the
                     CPU, oracle/glsl_ref/): G-buffer images, illumination and final frame of the same small fixed scenes.
  oracle_*.npz       oracle outputs for small fixed scenes (hit buffers, G-buffers, frame, counters).
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
uvt = importlib.import_module("unnamed-voxel-tracer_b200")
import oracle  # noqa: E402

ASSETS = "/root/reference/assets"
GOLD = os.path.join(ROOT, "tests", "golden")


def make_atlas():
    atlas = uvt.voxel.VoxelModelAtlas.init(None)
    for f in uvt.game.BLOCK_MODEL_FILES:
        atlas.load_block_model(os.path.join(ASSETS, f))
    models = atlas.models()
    np.save(os.path.join(GOLD, "atlas_models.npy"), models)
    counts = [int((m != 0).sum()) for m in models]
    json.dump({"files": uvt.game.BLOCK_MODEL_FILES, "filled": counts}, open(os.path.join(GOLD, "atlas_counts.json"), "w"), indent=1)
    return models, counts


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    models, counts = make_atlas()
    print(len(models), "models; filled:", counts)
    chicken = uvt.voxel.load_model(os.path.join(ASSETS, "chicken.vox"), 32)
    np.save(os.path.join(GOLD, "chicken_32.npy"), chicken)
    print("chicken.vox:", int((chicken != 0).sum()), "voxels of 32^3")


def make_oracle_goldens():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import camera_k0, camera_k1
    models = np.load(os.path.join(GOLD, "atlas_models.npy"))
    atlas = oracle.atlas_from_models(models)
    bm = uvt.voxel.VoxelBrickmap.init(512)
    uvt.procgen.procgen(512, bm)
    world = oracle.World(512, bm.chunks().copy(), bm.bricks().copy(), atlas)
    for name, cam in (("k0_96x54", camera_k0(oracle)), ("k1_96x54", camera_k1(uvt, oracle))):
        r = oracle.render(world, cam, 96, 54)
        pc = r["primary_counters"]
        np.savez_compressed(os.path.join(GOLD, f"oracle_w1_{name}.npz"), albedo=r["albedo"], normal=r["normal"],
                            position=r["position"], illumination=r["illumination"], frame=r["frame"], hits=r["hits"],
                            primary_counters=np.array([pc[k] for k in ("rays", "t_in", "t_chunk", "t_block", "hits")], dtype=np.uint64),
                            camera=np.frombuffer(cam.tobytes(), np.float32))
        print(name, pc, r["secondary_counters"])
        from oracle import glslref
        g = glslref.render(world, cam, 96, 54)
        np.savez_compressed(os.path.join(GOLD, f"glslref_w1_{name}.npz"), albedo=g["albedo"], normal=g["normal"], position=g["position"],
                            illumination=g["illumination"], frame=g["frame"], camera=np.frombuffer(cam.tobytes(), np.float32))


if __name__ == "__main__":
    make_oracle_goldens()
