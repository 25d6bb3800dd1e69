"""Host-side scene model above the C ABI (Python mirror used by tests, bench and tools).

Mirrors the reference's host classes for the hot path only:
  Scene  (reference src/scene/scene.{h,cpp})   -- id tables, dummies at index 0, rect/mesh lights
                                                   becoming emitter instances, shots, per-shot state
  Loader (reference src/loader/{loader,material}.cpp) -- the JSON scene format
  Camera (reference src/core/camera.cpp)        -- GpuCamera matrices
  EnvMap (reference src/core/texture.cpp:144-226) -- marginal / conditional inverse-CDF tables
and the offline driver loop of reference src/tracer/tracer.cpp:177-264 (`render_shot`).

All arithmetic that feeds the device is done in float32 in the same order as the reference
where the order matters (env-map prefix sums, diffuse Fresnel integral).
"""
import json
import math
import os

import numpy as np

from . import structs as S

f32 = np.float32
NV_TO_RAD = f32(math.pi / 180.0)


# ----------------------------------------------------------------------------- nvmath subset
def mat_identity():
    return np.eye(4, dtype=f32)


def translation(v):
    m = mat_identity()
    m[:3, 3] = v
    return m


def scaling(v):
    m = mat_identity()
    m[0, 0], m[1, 1], m[2, 2] = v
    return m


def rotation_x(a):
    c, s = f32(math.cos(a)), f32(math.sin(a))
    m = mat_identity()
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m


def rotation_y(a):
    c, s = f32(math.cos(a)), f32(math.sin(a))
    m = mat_identity()
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m


def rotation_z(a):
    c, s = f32(math.cos(a)), f32(math.sin(a))
    m = mat_identity()
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def _normalize(v):
    v = np.asarray(v, f32)
    return v / f32(np.sqrt(np.dot(v, v)))


def look_at(eye, center, up):
    """reference ext/nvpro_core/nvmath/nvmath.inl:979-1019"""
    eye, center, up = (np.asarray(a, f32) for a in (eye, center, up))
    z = _normalize(eye - center)
    x = np.cross(up, z).astype(f32)
    y = np.cross(z, x).astype(f32)
    x, y = _normalize(x), _normalize(y)
    m = mat_identity()
    m[0, :3], m[1, :3], m[2, :3] = x, y, z
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(x, eye), -np.dot(y, eye), -np.dot(z, eye)
    return m


def invert_rot_trans(a):
    """reference ext/nvpro_core/nvmath/nvmath.inl:911-931"""
    b = np.zeros((4, 4), f32)
    b[:3, :3] = a[:3, :3].T
    b[3, :3] = a[3, :3]
    b[:3, 3] = -(a[:3, :3].T @ a[:3, 3])
    b[3, 3] = a[3, 3]
    return b.astype(f32)


def colmajor(m):
    """4x4 row-indexed numpy matrix -> 16 floats in nvmath (column-major) order."""
    return np.ascontiguousarray(np.asarray(m, f32).T).reshape(16)


def from_json_mat4(v):
    """reference src/loader/utils.h:30-40 -- JSON matrices are row-major in the file."""
    return np.asarray(v, f32).reshape(4, 4)


# ----------------------------------------------------------------------------- camera
def perspective_raster_to_camera(width, height, fov_deg, near=0.1, far=100.0):
    """reference src/core/camera.cpp:28-62,93-94 (fov is horizontal)."""
    recip = 1.0 / (far - near)
    ctot = 1.0 / math.tan(math.radians(fov_deg) * 0.5)
    persp = np.zeros((4, 4), np.float64)
    persp[0, 0] = persp[1, 1] = ctot
    persp[2, 2] = far * recip
    persp[2, 3] = -near * far * recip
    persp[3, 2] = 1
    aspect = width / float(height)
    m = scaling((1, aspect, 1)).astype(np.float64) @ persp
    m = translation((1, 1, 0)).astype(np.float64) @ m
    m = scaling((0.5, 0.5, 1)).astype(np.float64) @ m
    m = scaling((width, height, 1)).astype(np.float64) @ m
    return np.linalg.inv(m).astype(f32)


class Shot:
    """reference src/core/camera.h:8-15"""

    def __init__(self, eye, lookat, up, env_transform=None, state=None):
        self.eye, self.lookat, self.up = (np.asarray(a, f32) for a in (eye, lookat, up))
        self.env_transform = mat_identity() if env_transform is None else np.asarray(env_transform, f32)
        self.state = state  # per-shot State override (or None -> scene default)


# ----------------------------------------------------------------------------- materials
COMPLEX_IOR = {  # reference src/loader/material.cpp:250-291 (eta, k per RGB)
    "a-C": ((2.9440999183, 2.2271502925, 1.9681668794), (0.8874329109, 0.7993216383, 0.8152862927)),
    "Ag": ((0.1552646489, 0.1167232965, 0.1383806959), (4.8283433224, 3.1222459278, 2.1469504455)),
    "Al": ((1.6574599595, 0.8803689579, 0.5212287346), (9.2238691996, 6.2695232477, 4.8370012281)),
    "AlAs": ((3.6051023902, 3.2329365777, 2.2175611545), (0.0006670247, -0.0004999400, 0.0074261204)),
    "AlSb": ((-0.0485225705, 4.1427547893, 4.6697691348), (-0.0363741915, 0.0937665154, 1.3007390124)),
    "Au": ((0.1431189557, 0.3749570432, 1.4424785571), (3.9831604247, 2.3857207478, 1.6032152899)),
    "Be": ((4.1850592788, 3.1850604423, 2.7840913457), (3.8354398268, 3.0101260162, 2.8690088743)),
    "Cr": ((4.3696828663, 2.9167024892, 1.6547005413), (5.2064337956, 4.2313645277, 3.7549467933)),
    "CsI": ((2.1449030413, 1.7023164587, 1.6624194173), (0.0, 0.0, 0.0)),
    "Cu": ((0.2004376970, 0.9240334304, 1.1022119527), (3.9129485033, 2.4528477015, 2.1421879552)),
    "Cu2O": ((3.5492833755, 2.9520622449, 2.7369202137), (0.1132179294, 0.1946659670, 0.6001681264)),
    "CuO": ((3.2453822204, 2.4496293965, 2.1974114493), (0.5202739621, 0.5707372756, 0.7172250613)),
    "d-C": ((2.7112524747, 2.3185812849, 2.2288565009), (0.0, 0.0, 0.0)),
    "Hg": ((2.3989314904, 1.4400254917, 0.9095512090), (6.3276269444, 4.3719414152, 3.4217899270)),
    "HgTe": ((4.7795267752, 3.2309984581, 2.6600252401), (1.6319827058, 1.5808189339, 1.7295753852)),
    "Ir": ((3.0864098394, 2.0821938440, 1.6178866805), (5.5921510077, 4.0671757150, 3.2672611269)),
    "K": ((0.0640493070, 0.0464100621, 0.0381842017), (2.1042155920, 1.3489364357, 0.9132113889)),
    "Li": ((0.2657871942, 0.1956102432, 0.2209198538), (3.5401743407, 2.3111306542, 1.6685930000)),
    "MgO": ((2.0895885542, 1.6507224525, 1.5948759692), (0.0, -0.0, 0.0)),
    "Mo": ((4.4837010280, 3.5254578255, 2.7760769438), (4.1111307988, 3.4208716252, 3.1506031404)),
    "Na": ((0.0602665320, 0.0561412435, 0.0619909494), (3.1792906496, 2.1124800781, 1.5790940266)),
    "Nb": ((3.4201353595, 2.7901921379, 2.3955856658), (3.4413817900, 2.7376437930, 2.5799132708)),
    "Ni": ((2.3672753521, 1.6633583302, 1.4670554172), (4.4988329911, 3.0501643957, 2.3454274399)),
    "Rh": ((2.5857954933, 1.8601866068, 1.5544279524), (6.7822927110, 4.7029501026, 3.9760892461)),
    "Se-e": ((5.7242724833, 4.1653992967, 4.0816099264), (0.8713747439, 1.1052845009, 1.5647788766)),
    "Se": ((4.0592611085, 2.8426947380, 2.8207582835), (0.7543791750, 0.6385150558, 0.5215872029)),
    "SiC": ((3.1723450205, 2.5259677964, 2.4793623897), (0.0000007284, -0.0000006859, 0.0000100150)),
    "SnTe": ((4.5251865890, 1.9811525984, 1.2816819226), (0.0, 0.0, 0.0)),
    "Ta": ((2.0625846607, 2.3930915569, 2.6280684948), (2.4080467973, 1.7413705864, 1.9470377016)),
    "Te-e": ((7.5090397678, 4.2964603080, 2.3698732430), (5.5842076830, 4.9476231084, 3.9975145063)),
    "Te": ((7.3908396088, 4.4821028985, 2.6370708478), (3.2561412892, 3.5273908133, 3.2921683116)),
    "ThF4": ((1.8307187117, 1.4422274283, 1.3876488528), (0.0, 0.0, 0.0)),
    "TiC": ((3.7004673762, 2.8374356509, 2.5823030278), (3.2656905818, 2.3515586388, 2.1727857800)),
    "TiN": ((1.6484691607, 1.1504482522, 1.3797795097), (3.3684596226, 1.9434888540, 1.1020123347)),
    "TiO2-e": ((3.1065574823, 2.5131551146, 2.5823844157), (0.0000289537, -0.0000251484, 0.0001775555)),
    "TiO2": ((3.4566203131, 2.8017076558, 2.9051485020), (0.0001026662, -0.0000897534, 0.0006356902)),
    "VC": ((3.6575665991, 2.7527298065, 2.5326814570), (3.0683516659, 2.1986687713, 1.9631816252)),
    "VN": ((2.8656011588, 2.1191817791, 1.9400767149), (3.0323264950, 2.0561075580, 1.6162930914)),
    "V": ((4.2775126218, 3.5131538236, 2.7611257461), (3.4911844504, 2.8893580874, 3.1116965117)),
    "W": ((4.3707029924, 3.3002972445, 2.9982666528), (3.5006778591, 2.6048652781, 2.2731930614)),
}


def _dielectric_reflectance(eta, cos_i):
    """reference src/loader/material.cpp:6-23 (float32)."""
    eta, cos_i = f32(eta), f32(cos_i)
    if cos_i < 0:
        eta, cos_i = f32(1) / eta, -cos_i
    sin2 = eta * eta * (f32(1) - cos_i * cos_i)
    if sin2 > 1:
        return f32(1)
    cos_t = f32(np.sqrt(max(f32(1) - sin2, f32(0))))
    rs = (eta * cos_i - cos_t) / (eta * cos_i + cos_t)
    rp = (eta * cos_t - cos_i) / (eta * cos_t + cos_i)
    return (rs * rs + rp * rp) * f32(0.5)


def compute_diffuse_fresnel(ior, n=1000):
    """reference src/loader/material.cpp:25-37 -- trapezoid rule, double accumulator."""
    acc = 0.0
    fb = _dielectric_reflectance(ior, 0.0)
    for i in range(1, n + 1):
        cos2 = f32(i) / f32(n)
        fa = _dielectric_reflectance(ior, min(f32(np.sqrt(cos2)), f32(1)))
        acc += float(f32(fa + fb)) * (0.5 / n)
        fb = fa
    return f32(acc)


_MATERIAL_TYPES = {
    "brdf_lambertian": S.MAT_LAMBERTIAN, "brdf_pbr_metalness_roughness": S.MAT_PBR, "brdf_emissive": S.MAT_EMISSIVE,
    "brdf_kang18": S.MAT_KANG18, "bsdf_dielectric": S.MAT_DIELECTRIC, "brdf_plastic": S.MAT_PLASTIC,
    "brdf_rough_plastic": S.MAT_ROUGH_PLASTIC, "brdf_conductor": S.MAT_CONDUCTOR,
    "brdf_mirror": S.MAT_MIRROR, "brdf_rough_conductor": S.MAT_ROUGH_CONDUCTOR, "brdf_disney": S.MAT_DISNEY,
    "brdf_phong": S.MAT_PHONG}
IN_SCOPE_MATERIALS = tuple(range(12))  # all twelve closest-hit shaders of the reference (src/shared/material.h:7-21)


# ----------------------------------------------------------------------------- env map
def envmap_tables(rgba):
    """reference src/core/texture.cpp:144-226.  Returns (marginal, conditional) RGBA32F tables."""
    img = np.ascontiguousarray(rgba, f32)
    h, w = img.shape[:2]
    weight = (0.3 * img[..., 0].astype(np.float64) + 0.6 * img[..., 1].astype(np.float64)
              + 0.1 * img[..., 2].astype(np.float64)).astype(f32)  # double arithmetic, texture.cpp:167
    cdf2d = np.cumsum(weight, axis=1, dtype=f32)            # sequential fp32 row prefix sums
    row_sum = cdf2d[:, -1].copy()
    denom = row_sum.astype(np.float64) + 1e-7
    pdf2d = (weight.astype(np.float64) / denom[:, None]).astype(f32)
    cdf2d = (cdf2d.astype(np.float64) / denom[:, None]).astype(f32)
    cdf1d = np.cumsum(row_sum, dtype=f32)
    total = float(cdf1d[-1]) + 1e-7
    pdf1d = (row_sum.astype(np.float64) / total).astype(f32)
    cdf1d = (cdf1d.astype(np.float64) / total).astype(f32)
    marginal = np.zeros((h, w, 4), f32)
    conditional = np.zeros((h, w, 4), f32)
    inv_h = (np.arange(1, h + 1, dtype=f32) / f32(h)).astype(f32)
    rows = np.searchsorted(cdf1d, inv_h, side="left").astype(f32)
    marginal[:, 0, 0] = rows / f32(h)
    marginal[:, 0, 1] = pdf1d
    inv_w = (np.arange(1, w + 1, dtype=f32) / f32(w)).astype(f32)
    for j in range(h):
        cols = np.searchsorted(cdf2d[j], inv_w, side="left").astype(f32)
        conditional[j, :, 0] = cols / f32(w)
    conditional[:, :, 1] = pdf2d
    return marginal, conditional


# ----------------------------------------------------------------------------- meshes
def make_vertices(pos, uv=None, normal=None):
    pos = np.asarray(pos, f32).reshape(-1, 3)
    v = np.zeros(pos.shape[0], S.Vertex)
    v["pos"] = pos
    if uv is not None:
        v["uv"] = np.asarray(uv, f32).reshape(-1, 2)
    if normal is not None:
        v["normal"] = np.asarray(normal, f32).reshape(-1, 3)
    return v


def load_obj(path):
    """reference src/core/mesh.cpp:112-145 semantics: triangulated fan, vertices unrolled per
    face corner (no index dedup), uv.y = 1 - v, indices = 0..n-1."""
    P, T, N = [], [], []
    corners = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                P.append([float(x) for x in t[1:4]])
            elif t[0] == "vt":
                T.append([float(x) for x in t[1:3]])
            elif t[0] == "vn":
                N.append([float(x) for x in t[1:4]])
            elif t[0] == "f":
                idx = []
                for c in t[1:]:
                    parts = c.split("/")
                    vi = int(parts[0])
                    ti = int(parts[1]) if len(parts) > 1 and parts[1] else 0
                    ni = int(parts[2]) if len(parts) > 2 and parts[2] else 0
                    idx.append((vi - 1 if vi > 0 else len(P) + vi, ti - 1 if ti > 0 else (len(T) + ti if ti < 0 else -1),
                                ni - 1 if ni > 0 else (len(N) + ni if ni < 0 else -1)))
                for k in range(1, len(idx) - 1):
                    corners += [idx[0], idx[k], idx[k + 1]]
    P, T, N = np.asarray(P, f32).reshape(-1, 3), np.asarray(T, f32).reshape(-1, 2), np.asarray(N, f32).reshape(-1, 3)
    c = np.asarray(corners, np.int64).reshape(-1, 3)
    v = np.zeros(c.shape[0], S.Vertex)
    v["pos"] = P[c[:, 0]]
    if len(T):
        has = c[:, 1] >= 0
        uv = T[np.where(has, c[:, 1], 0)]
        v["uv"][has] = np.stack([uv[:, 0], f32(1) - uv[:, 1]], 1)[has]
    if len(N):
        has = c[:, 2] >= 0
        v["normal"][has] = N[np.where(has, c[:, 2], 0)][has]
    return v, np.arange(c.shape[0], dtype=np.uint32)


def save_obj(path, vertices, indices):
    """Writes an indexed OBJ whose `load_obj` round trip reproduces `vertices[indices]` corner by corner."""
    v = np.asarray(vertices, S.Vertex)
    idx = np.asarray(indices, np.uint32).reshape(-1, 3)
    with open(path, "w") as f:
        for p in v["pos"]:
            f.write("v %.9g %.9g %.9g\n" % tuple(p))
        for t in v["uv"]:
            f.write("vt %.9g %.9g\n" % (t[0], 1.0 - float(t[1])))
        for n in v["normal"]:
            f.write("vn %.9g %.9g %.9g\n" % tuple(n))
        for a, b, c in idx + 1:
            f.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")


def load_image(path, gamma=1.0):
    """reference src/core/texture.cpp:306-339: LDR decoded with pow(x/255, gamma); .hdr/.npy linear."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        a = np.load(path).astype(f32)
    elif ext == ".hdr":
        import cv2
        a = cv2.imread(path, cv2.IMREAD_UNCHANGED)[..., ::-1].astype(f32)
    else:
        from PIL import Image
        im = np.asarray(Image.open(path).convert("RGBA"), f32) / f32(255)
        a = im.copy()
        a[..., :3] = np.power(im[..., :3], f32(gamma))
    if a.ndim == 2:
        a = np.repeat(a[..., None], 3, 2)
    if a.shape[2] == 3:
        a = np.concatenate([a, np.ones(a.shape[:2] + (1,), f32)], 2)
    return np.ascontiguousarray(a, f32)


# ----------------------------------------------------------------------------- scene
class Scene:
    def __init__(self):
        # reference src/scene/scene.cpp:87-130: dummies at index 0
        self.textures = [np.zeros((1, 1, 4), f32)]
        self.texture_ids = {"add_by_default_dummy_texture": 0}
        self.materials = [S.default_material()]
        self.material_ids = {"add_by_default_dummy_material": 0}
        self.lights = [S.dummy_light()]
        self.meshes = []  # (vertices, indices)
        self.mesh_ids = {}
        self.instances = []  # (xform 4x4, mesh, material, light)
        self.envmap = None  # (rgba, marginal, conditional)
        self.sunsky = S.default_sunsky()
        self.state = S.default_state()
        self.shots = []
        self.camera = None  # dict(type, width, height, fov | fxfycxcy, aperture, focal_distance)
        self.output = {"hdr": False, "render_result": True, "channel_ldr": []}
        self.base_dir = "."

    # ---- Scene::add*
    def add_texture(self, name, rgba):
        self.texture_ids[name] = len(self.textures)
        self.textures.append(np.ascontiguousarray(rgba, f32))
        return self.texture_ids[name]

    def add_material(self, name, material):
        self.material_ids[name] = len(self.materials)
        self.materials.append(np.array(material, S.Material))
        return self.material_ids[name]

    def add_mesh(self, name, vertices, indices):
        self.mesh_ids[name] = len(self.meshes)
        self.meshes.append((np.ascontiguousarray(vertices, S.Vertex), np.ascontiguousarray(indices, np.uint32).reshape(-1)))
        return self.mesh_ids[name]

    def add_instance(self, mesh, material, xform=None):
        mesh = self.mesh_ids[mesh] if isinstance(mesh, str) else mesh
        material = self.material_ids[material] if isinstance(material, str) else material
        self.instances.append((mat_identity() if xform is None else np.asarray(xform, f32), mesh, material, -1))

    def add_light(self, light):
        """reference src/scene/scene.cpp:220-246: rect lights also become a 2-triangle emitter instance."""
        light = np.array(light, S.Light)
        light_id = len(self.lights)
        if light["type"] == S.LIGHT_RECT:
            p, u, v = light["position"], light["u"], light["v"]
            verts = make_vertices([p, p + u, p + u + v, p + v])  # reference src/core/mesh.cpp:19-32
            mid = self.add_mesh(f"__rectLight:{light_id}", verts, [0, 1, 2, 0, 2, 3])
            self.instances.append((mat_identity(), mid, 0, light_id))
        self.lights.append(light)
        return light_id

    def add_mesh_light(self, radiance, vertices, indices):
        """reference src/scene/scene.cpp:248-283: one triangle light + emitter instance per facet."""
        v = np.asarray(vertices, S.Vertex)
        idx = np.asarray(indices, np.uint32).reshape(-1, 3)
        for a, b, c in idx:
            l = np.zeros((), S.Light)
            l["type"] = S.LIGHT_TRIANGLE
            l["radiance"] = radiance
            l["position"] = v["pos"][a]
            l["u"] = v["pos"][b] - v["pos"][a]
            l["v"] = v["pos"][c] - v["pos"][a]
            l["area"] = f32(np.linalg.norm(np.cross(l["u"], l["v"]))) * f32(0.5)
            light_id = len(self.lights)
            verts = make_vertices([l["position"], l["position"] + l["u"], l["position"] + l["v"]])
            mid = self.add_mesh(f"__meshLight:{light_id}", verts, [0, 1, 2])
            self.instances.append((mat_identity(), mid, 0, light_id))
            self.lights.append(l)

    def set_envmap(self, rgba):
        rgba = np.ascontiguousarray(rgba, f32)
        marginal, conditional = envmap_tables(rgba)
        self.envmap = (rgba, marginal, conditional)
        self.state["hasEnvMap"] = 1
        self.state["envMapResolution"] = (rgba.shape[1], rgba.shape[0])

    def set_camera(self, type="perspective", width=512, height=512, fov=45.0, aperture=0.0, focal_distance=0.1,
                   fxfycxcy=None):
        self.camera = dict(type=type, width=int(width), height=int(height), fov=float(fov), aperture=float(aperture),
                           focal_distance=float(focal_distance), fxfycxcy=fxfycxcy)

    def set_channels(self, names):
        """reference src/loader/loader.cpp:163-185"""
        st = self.state
        st["nMultiChannel"] = len(names)
        for cid, n in enumerate(names):
            st[n + "OutChannel"] = cid
        self.output["channel_ldr"] = [False] * len(names)

    # ---- GPU structs
    def gpu_camera(self, shot):
        """reference src/core/camera.cpp:77-99 with view = scale(1,-1,-1) * look_at (camera.h:26-28)."""
        c = np.zeros((), S.Camera)
        cam = self.camera
        view = scaling((1, -1, -1)) @ look_at(shot.eye, shot.lookat, shot.up)
        c["cameraToWorld"] = colmajor(invert_rot_trans(view.astype(f32)))
        c["envTransform"] = colmajor(shot.env_transform)
        if cam["type"] == "perspective":
            c["type"] = S.CAMERA_PERSPECTIVE
            fov = min(max(cam["fov"], 0.01), 179.0)  # cameramanipulator.cpp:321-324
            c["rasterToCamera"] = colmajor(perspective_raster_to_camera(cam["width"], cam["height"], fov))
            c["focalDistance"] = cam["focal_distance"]
            c["aperture"] = cam["aperture"]
        else:
            c["type"] = S.CAMERA_OPENCV
            c["fxfycxcy"] = cam["fxfycxcy"]
        return c

    def shot_state(self, shot_id):
        """reference src/scene/scene.cpp:439-453: only six fields are taken from the shot's state."""
        st = self.state.copy()
        sh = self.shots[shot_id].state
        if sh is not None:
            for k in ("spp", "maxPathDepth", "useFaceNormal", "ignoreEmissive", "envMapIntensity", "bgColor"):
                st[k] = sh[k]
        st["numLights"] = len(self.lights) - 1  # scene.cpp:33
        return st

    # ---- upload through the C ABI (≙ Scene::submit + PipelineRaytrace::init)
    def upload(self, ctx):
        ctx.set_film(self.camera["width"], self.camera["height"])
        for t in self.textures:
            ctx.add_texture(t)
        for m in self.materials:
            ctx.add_material(m)
        ctx.set_lights(np.array(self.lights, S.Light))
        if self.envmap is not None:
            ctx.set_envmap(*self.envmap)
        for v, i in self.meshes:
            ctx.add_mesh(v, i)
        for x, mesh, mat, light in self.instances:
            ctx.add_instance(colmajor(x), mesh, mat, light)
        ctx.set_sunsky(self.sunsky)
        return ctx.build_accel()

    def begin_shot(self, ctx, shot_id):
        """≙ Scene::setShot + setSpp(1) + resetFrame (reference src/tracer/tracer.cpp:206-212). Returns total spp."""
        st = self.shot_state(shot_id)
        tot = int(st["spp"])
        st["spp"] = 1
        st["curFrame"] = -1
        ctx.set_camera(self.gpu_camera(self.shots[shot_id]))
        ctx.set_sunsky(self.sunsky)
        ctx.set_state(st)
        ctx.reset_frame()
        return tot

    def render_shot(self, ctx, shot_id, spp=None):
        """≙ the per-shot loop of Tracer::runOffline (reference src/tracer/tracer.cpp:203-261)."""
        tot = self.begin_shot(ctx, shot_id)
        ctx.render_frames(tot if spp is None else spp)
        ctx.sync()
        n = int(self.state["nMultiChannel"])
        return [ctx.read_channel(0)] + [ctx.read_channel(1 + c) for c in range(n)]


# ----------------------------------------------------------------------------- JSON loader
def _parse_state(js, st, output):
    """reference src/loader/loader.cpp:148-234"""
    pt = js.get("path_tracing", {})
    if "spp" in pt:
        st["spp"] = pt["spp"]
    if "max_path_depth" in pt:
        st["maxPathDepth"] = pt["max_path_depth"]
    if "use_face_normal" in pt:
        st["useFaceNormal"] = 1 if pt["use_face_normal"] else 0
    if "ignore_emissive" in pt:
        st["ignoreEmissive"] = 1 if pt["ignore_emissive"] else 0
    if "background_color" in pt:
        st["bgColor"] = pt["background_color"]
    if "envmap_intensity" in pt:
        st["envMapIntensity"] = pt["envmap_intensity"]
    if "multi_channel" in pt:
        names = pt["multi_channel"]
        if len(names) > S.NUM_OUTPUT_IMAGES - 1:
            raise ValueError("channel numbers can not exceed 8")
        st["nMultiChannel"] = len(names)
        for cid, n in enumerate(names):
            if n in S.CHANNEL_NAMES:
                st[n + "OutChannel"] = cid
    if output is not None:
        if "path_tracing" in js:
            output["channel_ldr"] = [False] * int(st["nMultiChannel"])
            for cid, v in enumerate(pt.get("multi_channel_ldr", [])[: int(st["nMultiChannel"])]):
                output["channel_ldr"][cid] = bool(v)
        if "output_render_result" in js:
            output["render_result"] = bool(js["output_render_result"])
        if "output_hdr" in js:
            output["hdr"] = bool(js["output_hdr"])


def _parse_toworld(js, ban_translation=False):
    """reference src/loader/loader.cpp:366-396"""
    m = mat_identity()
    for s in js:
        t, v = s["type"], s["value"]
        if t == "matrix":
            x = from_json_mat4(v)
        elif t == "translate" and not ban_translation:
            x = translation(v)
        elif t == "scale":
            x = scaling(v)
        elif t == "rotx":
            x = rotation_x(NV_TO_RAD * f32(v))
        elif t == "roty":
            x = rotation_y(NV_TO_RAD * f32(v))
        elif t == "rotz":
            x = rotation_z(NV_TO_RAD * f32(v))
        elif t == "rotate":
            x = rotation_z(NV_TO_RAD * f32(v[2])) @ rotation_y(NV_TO_RAD * f32(v[1])) @ rotation_x(NV_TO_RAD * f32(v[0]))
        else:
            raise ValueError(f"unrecognized toworld singleton type [{t}]")
        m = (x @ m).astype(f32)
    return m


def _parse_material(scene, js):
    """reference src/loader/material.cpp:54-405"""
    m = S.default_material()
    t = js["type"]
    if t not in _MATERIAL_TYPES:
        raise ValueError(f"unrecognized material type [{t}]")
    m["type"] = _MATERIAL_TYPES[t]
    tex = lambda key: scene.texture_ids[js[key]]

    def opt(key, field, conv=lambda x: x):
        if key in js:
            m[field] = conv(js[key])

    def opt_tex(key, field):
        if key in js:
            m[field] = tex(key)

    if t in ("brdf_lambertian", "brdf_mirror"):
        opt("diffuse_reflectance", "diffuse"), opt_tex("diffuse_texture", "diffuseTextureId")
        opt_tex("normal_texture", "normalTextureId")
    elif t == "brdf_pbr_metalness_roughness":
        opt_tex("normal_texture", "normalTextureId"), opt("diffuse_reflectance", "diffuse")
        opt_tex("diffuse_texture", "diffuseTextureId"), opt("metalness", "metalness")
        opt_tex("metalness_texture", "metalnessTextureId"), opt("roughness", "roughness")
        opt_tex("roughness_texture", "roughnessTextureId")
        m["specular"] = 0.0  # used as opacity
        opt_tex("opacity_texture", "opacityTextureId")
    elif t == "brdf_emissive":
        opt("radiance", "radiance"), opt("radiance_factor", "radianceFactor"), opt_tex("radiance_texture", "radianceTextureId")
    elif t == "brdf_kang18":
        opt_tex("normal_texture", "normalTextureId"), opt_tex("tangent_texture", "tangentTextureId")
        if "diffuse_texture" in js:
            m["diffuseTextureId"] = tex("diffuse_texture")
        else:
            m["diffuse"] = js["diffuse_reflectance"]
        if "specular_texture" in js:
            m["metalnessTextureId"] = tex("specular_texture")
        else:
            m["rhoSpec"] = js["specular_reflectance"]
        if "alpha_texture" in js:
            m["roughnessTextureId"] = tex("alpha_texture")
        else:
            m["anisoAlpha"] = js["alpha"]
        m["metalness"] = 0.0  # used as opacity
        opt_tex("opacity_texture", "opacityTextureId")
    elif t == "bsdf_dielectric":
        opt_tex("normal_texture", "normalTextureId"), opt("ior", "ior")
    elif t in ("brdf_plastic", "brdf_rough_plastic"):
        opt("ior", "ior"), opt("diffuse_reflectance", "diffuse"), opt_tex("diffuse_texture", "diffuseTextureId")
        opt_tex("normal_texture", "normalTextureId")
        if t == "brdf_rough_plastic":
            opt("alpha", "anisoAlpha"), opt_tex("alpha_texture", "roughnessTextureId")
        m["radiance"][0] = compute_diffuse_fresnel(float(m["ior"]), 1000)
    elif t in ("brdf_conductor", "brdf_rough_conductor"):
        name = js.get("material", "Cu")
        if name not in COMPLEX_IOR:
            raise ValueError(f"unrecognized material name in brdf_conductor [{name}]")
        m["radiance"], m["radianceFactor"] = COMPLEX_IOR[name]
        opt("diffuse_reflectance", "diffuse"), opt_tex("diffuse_texture", "diffuseTextureId")
        opt_tex("normal_texture", "normalTextureId")
        if t == "brdf_rough_conductor":
            opt("alpha", "anisoAlpha"), opt_tex("alpha_texture", "roughnessTextureId")
    elif t == "brdf_disney":
        opt_tex("normal_texture", "normalTextureId"), opt("diffuse_reflectance", "diffuse")
        opt_tex("diffuse_texture", "diffuseTextureId"), opt("metallic", "metalness")
        opt_tex("metallic_texture", "metalnessTextureId"), opt("roughness", "roughness")
        opt_tex("roughness_texture", "roughnessTextureId")
        m["rhoSpec"][0] = js.get("opacity", 0.0)
        opt_tex("opacity_texture", "opacityTextureId")
    elif t == "brdf_phong":
        opt_tex("normal_texture", "normalTextureId"), opt("diffuse_reflectance", "diffuse")
        opt_tex("diffuse_texture", "diffuseTextureId"), opt("specular_reflectance", "rhoSpec"), opt("shininess", "specular")
    return m


def load_scene_json(path):
    """reference src/loader/loader.cpp:67-144 -- parse order is significant (ids by insertion)."""
    with open(path) as f:
        js = json.load(f)
    for k in ("state", "camera", "meshes", "instances"):
        if k not in js:
            raise ValueError(f'missing key ["{k}"]')
    sc = Scene()
    sc.base_dir = os.path.dirname(os.path.abspath(path))
    rel = lambda p: p if os.path.isabs(p) else os.path.join(sc.base_dir, p)
    _parse_state(js["state"], sc.state, sc.output)
    cj = js["camera"]
    w, h = int(cj["film"]["resolution"][0]), int(cj["film"]["resolution"][1])
    if cj["type"] == "perspective":
        sc.set_camera("perspective", w, h, cj.get("fov", 45.0), cj.get("aperture", 0.0), cj.get("focal_distance", 0.1))
    elif cj["type"] == "opencv":
        sc.set_camera("opencv", w, h, fxfycxcy=[cj["fx"], cj["fy"], cj["cx"], cj["cy"]])
    else:
        raise ValueError(f"unrecognized camera type [{cj['type']}]")
    for tj in js.get("textures", []):
        sc.add_texture(tj["name"], load_image(rel(tj["path"]), tj.get("gamma", 1.0)))
    for mj in js.get("materials", []):
        sc.add_material(mj["name"], _parse_material(sc, mj))
    for lj in js.get("lights", []):
        l = np.zeros((), S.Light)
        l["type"] = S.LIGHT_UNDEFINED
        l["radiance"] = lj["radiance"]
        t = lj["type"]
        if t in ("rect", "triangle"):  # A.3-3: "triangle" is stored as a rect with halved area
            l["position"] = lj["position"]
            l["u"] = np.asarray(lj["v1"], f32) - l["position"]
            l["v"] = np.asarray(lj["v2"], f32) - l["position"]
            l["area"] = f32(np.linalg.norm(np.cross(l["u"], l["v"]).astype(f32)))
            l["type"] = S.LIGHT_RECT
            l["doubleSide"] = 1 if lj.get("double_side", False) else 0
            if t == "triangle":
                l["area"] *= f32(0.5)
        elif t == "point":
            l["position"] = lj["position"]
            l["type"] = S.LIGHT_POINT
        elif t == "distant":
            l["direction"] = lj["direction"]
            l["type"] = S.LIGHT_DIRECTIONAL
        elif t == "mesh":
            v, i = load_obj(rel(lj["path"]))
            sc.add_mesh_light(lj["radiance"], v, i)
            continue
        else:
            raise ValueError(f"unrecognized light type [{t}]")
        sc.add_light(l)
    if "envmap" in js:
        sc.set_envmap(load_image(rel(js["envmap"]["path"])))
    if "sunsky" in js:  # documented extension: the reference can only enable it from the GUI
        for k, v in js["sunsky"].items():
            sc.sunsky[k] = v
        sc.sunsky["in_use"] = 1 if js["sunsky"].get("in_use", 1) else 0
    for mj in js["meshes"]:
        v, i = load_obj(rel(mj["path"]))
        if mj.get("recompute_normal", False):  # reference src/core/mesh.cpp:61-67
            p = v["pos"].reshape(-1, 3, 3)
            n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]).astype(f32)
            n /= np.linalg.norm(n, axis=1, keepdims=True).astype(f32)
            v["normal"] = np.repeat(n, 3, 0)
        if "uv_scale" in mj:
            v["uv"] *= np.asarray(mj["uv_scale"], f32)
        sc.add_mesh(mj["name"], v, i)
    for ij in js["instances"]:
        if "material" not in ij:
            raise ValueError("instance without material (reference dereferences a null default, A.3-13)")
        sc.add_instance(ij["mesh"], ij["material"], _parse_toworld(ij["toworld"]) if "toworld" in ij else None)
    for sj in js.get("shots", []):
        t = sj["type"]
        if t == "lookat":
            eye, lookat, up = sj["eye"], sj["lookat"], sj["up"]
        elif t == "toworld":  # loader.cpp:420-427
            c2w = from_json_mat4(sj["matrix"])
            eye, up, lookat = c2w[:3, 3], (c2w @ np.array([0, 1, 0, 0], f32))[:3], (c2w @ np.array([0, 0, 1, 1], f32))[:3]
        elif t == "opencv":  # loader.cpp:428-435
            c2w = invert_rot_trans(from_json_mat4(sj["matrix"]))
            eye, up, lookat = c2w[:3, 3], (c2w @ np.array([0, -1, 0, 0], f32))[:3], (c2w @ np.array([0, 0, 1, 1], f32))[:3]
        else:
            raise ValueError(f"unrecognized shot type [{t}]")
        st = None
        if "state" in sj:
            st = sc.state.copy()
            _parse_state(sj["state"], st, None)
        env = _parse_toworld(sj["env_toworld"], True) if "env_toworld" in sj else None
        sc.shots.append(Shot(eye, lookat, up, env, st))
    return sc
