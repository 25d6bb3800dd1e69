"""Writes a host.Scene as the reference's on-disk format: scene JSON + OBJ
meshes + texture / env-map images, so the same synthetic configs can be loaded by the C++ host CLI.

    python tools/gen_scenes.py <outdir> [cornell|glass|pbr|field|raybench|materials ...]

Textures and env maps are stored as float32 .npy (lossless; the loaders of this repo accept .npy next
to the reference's hdr/exr/png/jpg).  Emitter meshes/instances created by lights are not written: the
loader recreates them from the "lights" block, like the reference does (scene.cpp:220-283).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from asuna_b200 import host, scenes, structs as S  # noqa: E402

_TYPE_NAMES = {v: k for k, v in host._MATERIAL_TYPES.items()}


def _material_json(sc, name, m):
    tex_name = {v: k for k, v in sc.texture_ids.items()}
    t = int(m["type"])
    js = {"type": _TYPE_NAMES[t], "name": name}
    v3 = lambda k: [float(x) for x in m[k]]

    def tex(field, key):
        if int(m[field]) >= 0:
            js[key] = tex_name[int(m[field])]

    if t in (S.MAT_LAMBERTIAN, S.MAT_MIRROR):
        js["diffuse_reflectance"] = v3("diffuse")
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
    elif t == S.MAT_PBR:
        js.update(diffuse_reflectance=v3("diffuse"), metalness=float(m["metalness"]), roughness=float(m["roughness"]))
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
        tex("metalnessTextureId", "metalness_texture"), tex("roughnessTextureId", "roughness_texture")
        tex("opacityTextureId", "opacity_texture")
    elif t == S.MAT_EMISSIVE:
        js.update(radiance=v3("radiance"), radiance_factor=v3("radianceFactor"))
        tex("radianceTextureId", "radiance_texture")
    elif t == S.MAT_KANG18:
        js.update(diffuse_reflectance=v3("diffuse"), specular_reflectance=v3("rhoSpec"), alpha=[float(x) for x in m["anisoAlpha"]])
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
        tex("tangentTextureId", "tangent_texture"), tex("metalnessTextureId", "specular_texture")
        tex("roughnessTextureId", "alpha_texture"), tex("opacityTextureId", "opacity_texture")
    elif t == S.MAT_DIELECTRIC:
        js["ior"] = float(m["ior"])
        tex("normalTextureId", "normal_texture")
    elif t in (S.MAT_PLASTIC, S.MAT_ROUGH_PLASTIC):
        js.update(ior=float(m["ior"]), diffuse_reflectance=v3("diffuse"))
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
        if t == S.MAT_ROUGH_PLASTIC:
            js["alpha"] = [float(x) for x in m["anisoAlpha"]]
            tex("roughnessTextureId", "alpha_texture")
    elif t == S.MAT_DISNEY:
        js.update(diffuse_reflectance=v3("diffuse"), metallic=float(m["metalness"]), roughness=float(m["roughness"]),
                  opacity=float(m["rhoSpec"][0]))
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
        tex("metalnessTextureId", "metallic_texture"), tex("roughnessTextureId", "roughness_texture")
        tex("opacityTextureId", "opacity_texture")
    elif t == S.MAT_PHONG:
        js.update(diffuse_reflectance=v3("diffuse"), specular_reflectance=v3("rhoSpec"), shininess=float(m["specular"]))
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
    elif t in (S.MAT_CONDUCTOR, S.MAT_ROUGH_CONDUCTOR):
        if t == S.MAT_ROUGH_CONDUCTOR:
            js["alpha"] = [float(x) for x in m["anisoAlpha"]]
            tex("roughnessTextureId", "alpha_texture")
        for metal, (eta, k) in host.COMPLEX_IOR.items():
            if np.allclose(eta, m["radiance"], atol=1e-6) and np.allclose(k, m["radianceFactor"], atol=1e-6):
                js["material"] = metal
        js["diffuse_reflectance"] = v3("diffuse")
        tex("diffuseTextureId", "diffuse_texture"), tex("normalTextureId", "normal_texture")
    return js


def write_scene(sc, outdir, name):
    os.makedirs(outdir, exist_ok=True)
    cam, st = sc.camera, sc.state
    chan = [None] * int(st["nMultiChannel"])
    for n in S.CHANNEL_NAMES:
        c = int(st[n + "OutChannel"])
        if 0 <= c < len(chan):
            chan[c] = n
    js = {"state": {"path_tracing": {"spp": int(st["spp"]), "max_path_depth": int(st["maxPathDepth"]),
                                     "use_face_normal": bool(st["useFaceNormal"]), "ignore_emissive": bool(st["ignoreEmissive"]),
                                     "background_color": [float(x) for x in st["bgColor"]],
                                     "envmap_intensity": float(st["envMapIntensity"]), "multi_channel": chan},
                    "output_hdr": True}}
    cj = {"type": cam["type"], "film": {"resolution": [cam["width"], cam["height"]]}}
    if cam["type"] == "perspective":
        cj.update(fov=cam["fov"], aperture=cam["aperture"], focal_distance=cam["focal_distance"])
    else:
        cj.update(zip(("fx", "fy", "cx", "cy"), [float(x) for x in cam["fxfycxcy"]]))
    js["camera"] = cj
    js["textures"] = []
    for tname, tid in sc.texture_ids.items():
        if tid == 0:
            continue
        fn = f"{name}_tex_{tname}.npy"
        np.save(os.path.join(outdir, fn), sc.textures[tid])
        js["textures"].append({"name": tname, "path": fn})
    js["materials"] = [_material_json(sc, n, sc.materials[i]) for n, i in sc.material_ids.items() if i != 0]
    js["lights"] = []
    mesh_light_tris = []
    for l in sc.lights[1:]:
        t = int(l["type"])
        rad = [float(x) for x in l["radiance"]]
        if t == S.LIGHT_RECT:
            p = l["position"]
            js["lights"].append({"type": "rect", "radiance": rad, "position": [float(x) for x in p],
                                 "v1": [float(x) for x in p + l["u"]], "v2": [float(x) for x in p + l["v"]],
                                 "double_side": bool(l["doubleSide"])})
        elif t == S.LIGHT_POINT:
            js["lights"].append({"type": "point", "radiance": rad, "position": [float(x) for x in l["position"]]})
        elif t == S.LIGHT_DIRECTIONAL:
            js["lights"].append({"type": "distant", "radiance": rad, "direction": [float(x) for x in l["direction"]]})
        elif t == S.LIGHT_TRIANGLE:
            mesh_light_tris.append((rad, l["position"], l["position"] + l["u"], l["position"] + l["v"]))
    if mesh_light_tris:  # one "mesh" light per run of equal radiance
        rad = mesh_light_tris[0][0]
        pos = np.array([p for tri in mesh_light_tris for p in tri[1:]], np.float32)
        fn = f"{name}_meshlight.obj"
        host.save_obj(os.path.join(outdir, fn), host.make_vertices(pos), np.arange(len(pos), dtype=np.uint32))
        js["lights"].append({"type": "mesh", "radiance": rad, "path": fn})
    if sc.envmap is not None:
        fn = f"{name}_env.npy"
        np.save(os.path.join(outdir, fn), sc.envmap[0])
        js["envmap"] = {"path": fn}
    if int(sc.sunsky["in_use"]) == 1:
        js["sunsky"] = {k: (sc.sunsky[k].tolist() if sc.sunsky[k].ndim else sc.sunsky[k].item()) for k in S.SunSky.names}
    js["meshes"] = []
    for mname, mid in sc.mesh_ids.items():
        if mname.startswith("__"):
            continue
        fn = f"{name}_{mname}.obj"
        host.save_obj(os.path.join(outdir, fn), *sc.meshes[mid])
        js["meshes"].append({"name": mname, "path": fn})
    mesh_name = {v: k for k, v in sc.mesh_ids.items()}
    mat_name = {v: k for k, v in sc.material_ids.items()}
    js["instances"] = [{"mesh": mesh_name[m], "material": mat_name[a], "toworld": [{"type": "matrix", "value": [float(v) for v in x.reshape(-1)]}]}
                       for x, m, a, l in sc.instances if l < 0]
    js["shots"] = [{"type": "lookat", "eye": [float(v) for v in s.eye], "lookat": [float(v) for v in s.lookat],
                    "up": [float(v) for v in s.up]} for s in sc.shots]
    path = os.path.join(outdir, name + ".json")
    with open(path, "w") as f:
        json.dump(js, f, indent=1)
    return path


BUILDERS = {"cornell": scenes.cornell, "glass": scenes.glass_blob, "pbr": scenes.pbr_spheres,
            "field": scenes.instanced_field, "raybench": scenes.ray_bench,
            "materials": lambda: scenes.cornell_materials(env=True, lights="all", textured=True)}

if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "scenes_out"
    for n in (sys.argv[2:] or ["cornell"]):
        print(write_scene(BUILDERS[n](), out, n))
