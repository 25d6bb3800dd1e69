"""numpy mirrors of the wire structs in include/asuna_b200.h (reference src/shared/*.h).

Scalar layout, 4-byte words, no padding; sizes are asserted against the library at load
time (capi.Library.check_abi) so a drift between header and binding fails loudly.
"""
import numpy as np

f4, i4, u4 = np.float32, np.int32, np.uint32

Vertex = np.dtype([("pos", f4, 3), ("uv", f4, 2), ("normal", f4, 3), ("tangent", f4, 3)])  # vertex.h:6-11

Material = np.dtype([  # material.h:24-49
    ("diffuse", f4, 3), ("rhoSpec", f4, 3), ("anisoAlpha", f4, 2), ("ior", f4), ("roughness", f4),
    ("subsurface", f4), ("specular", f4), ("specularTint", f4), ("anisotropic", f4), ("sheen", f4),
    ("sheenTint", f4), ("clearcoat", f4), ("clearcoatGloss", f4), ("radiance", f4, 3), ("metalness", f4),
    ("radianceFactor", f4, 3), ("diffuseTextureId", i4), ("roughnessTextureId", i4),
    ("metalnessTextureId", i4), ("radianceTextureId", i4), ("normalTextureId", i4),
    ("tangentTextureId", i4), ("opacityTextureId", i4), ("type", u4)])

Light = np.dtype([  # light.h:16-25
    ("type", i4), ("position", f4, 3), ("direction", f4, 3), ("radiance", f4, 3), ("u", f4, 3),
    ("v", f4, 3), ("radius", f4), ("area", f4), ("doubleSide", u4)])

Camera = np.dtype([  # camera.h:15-24 (matrices column-major)
    ("rasterToCamera", f4, 16), ("cameraToWorld", f4, 16), ("envTransform", f4, 16),
    ("fxfycxcy", f4, 4), ("type", u4), ("aperture", f4), ("focalDistance", f4), ("padding", f4)])

State = np.dtype([  # pushconstant.h:10-34
    ("spp", i4), ("curFrame", i4), ("maxPathDepth", i4), ("numLights", i4), ("bgColor", f4, 3),
    ("useFaceNormal", u4), ("ignoreEmissive", u4), ("hasEnvMap", u4), ("envMapResolution", f4, 2),
    ("envMapIntensity", f4), ("nMultiChannel", u4), ("diffuseOutChannel", i4),
    ("specularOutChannel", i4), ("roughnessOutChannel", i4), ("normalOutChannel", i4),
    ("positionOutChannel", i4), ("tangentOutChannel", i4), ("uvOutChannel", i4)])

SunSky = np.dtype([  # sun_and_sky.h:6-28
    ("rgb_unit_conversion", f4, 3), ("multiplier", f4), ("haze", f4), ("redblueshift", f4),
    ("saturation", f4), ("horizon_height", f4), ("ground_color", f4, 3), ("horizon_blur", f4),
    ("night_color", f4, 3), ("sun_disk_intensity", f4), ("sun_direction", f4, 3),
    ("sun_disk_scale", f4), ("sun_glow_intensity", f4), ("y_is_up", i4),
    ("physically_scaled_sun", i4), ("in_use", i4)])

Stats = np.dtype([
    ("paths", np.uint64), ("closest_rays", np.uint64), ("shadow_rays", np.uint64),
    ("incoherent_closest_rays", np.uint64), ("kernel_launches", np.uint64), ("closest_launches", np.uint64),
    ("node_visits", np.uint64), ("tri_tests", np.uint64), ("trace_ms", f4), ("shade_ms", f4), ("total_ms", f4),
    ("build_ms", f4), ("closest_ms", f4), ("shadow_ms", f4), ("pad", f4, 2)])

# GpuPushConstantPost, reference src/shared/pushconstant.h:49-62
Post = np.dtype([("brightness", f4), ("contrast", f4), ("saturation", f4), ("vignette", f4), ("avgLum", f4), ("zoom", f4),
                 ("renderingRatio", f4, 2), ("autoExposure", np.int32), ("Ywhite", f4), ("key", f4), ("tmType", np.uint32)])
assert Post.itemsize == 48
TONE_MAPPERS = {"none": 0, "gamma": 1, "reinhard": 2, "Aces": 3, "filmic": 4, "pbrt": 5, "custom": 6}  # loader.cpp:205-222


def default_post(tone_mapping="filmic"):
    """reference src/core/state.h:46-58"""
    p = np.zeros((), Post)
    p["brightness"] = p["contrast"] = p["saturation"] = p["avgLum"] = p["zoom"] = 1.0
    p["renderingRatio"] = (1.0, 1.0)
    p["Ywhite"] = p["key"] = 0.5
    p["tmType"] = TONE_MAPPERS[tone_mapping]
    return p


EXPECTED_SIZES = (44, 132, 76, 224, 84, 96)
assert (Vertex.itemsize, Material.itemsize, Light.itemsize, Camera.itemsize, State.itemsize,
        SunSky.itemsize) == EXPECTED_SIZES

# material.h:7-21
MAT_LAMBERTIAN, MAT_KANG18, MAT_EMISSIVE, MAT_PBR, MAT_PLASTIC, MAT_ROUGH_PLASTIC = 0, 1, 2, 3, 4, 5
MAT_CONDUCTOR, MAT_ROUGH_CONDUCTOR, MAT_MIRROR, MAT_DISNEY, MAT_DIELECTRIC, MAT_PHONG = 6, 7, 8, 9, 10, 11
# light.h:7-13
LIGHT_DIRECTIONAL, LIGHT_RECT, LIGHT_TRIANGLE, LIGHT_POINT, LIGHT_UNDEFINED = 0, 1, 2, 3, 4
CAMERA_PERSPECTIVE, CAMERA_OPENCV = 0, 1
NUM_OUTPUT_IMAGES = 9

CHANNEL_NAMES = ("diffuse", "normal", "specular", "tangent", "roughness", "position", "uv")  # loader.cpp:177-183


def default_material():
    """reference src/core/material.h:9-35"""
    m = np.zeros((), Material)
    m["ior"] = 1.5
    m["roughness"] = 0.5
    m["radianceFactor"] = 1.0
    for k in ("diffuseTextureId", "radianceTextureId", "metalnessTextureId", "normalTextureId",
              "roughnessTextureId", "tangentTextureId", "opacityTextureId"):
        m[k] = -1
    m["type"] = MAT_LAMBERTIAN
    return m


def default_state():
    """reference src/core/state.h:15-44"""
    s = np.zeros((), State)
    s["curFrame"] = -1
    s["spp"] = 1
    s["maxPathDepth"] = 3
    s["envMapIntensity"] = 1.0
    for k in ("diffuseOutChannel", "specularOutChannel", "roughnessOutChannel", "normalOutChannel",
              "positionOutChannel", "tangentOutChannel", "uvOutChannel"):
        s[k] = -1
    return s


def default_sunsky():
    """reference src/scene/scene.cpp:112-129"""
    s = np.zeros((), SunSky)
    s["rgb_unit_conversion"] = 1.0
    s["multiplier"] = 0.0000101320
    s["saturation"] = 1.0
    s["ground_color"] = 0.4
    s["horizon_blur"] = 0.1
    s["night_color"] = (0.0, 0.0, 0.01)
    s["sun_disk_intensity"] = 0.8
    s["sun_direction"] = (0.0, 0.78, 0.62)
    s["sun_disk_scale"] = 5.0
    s["sun_glow_intensity"] = 1.0
    s["y_is_up"] = 1
    s["physically_scaled_sun"] = 1
    s["in_use"] = 0
    return s


def dummy_light():
    """reference src/scene/scene.cpp:100-111"""
    l = np.zeros((), Light)
    l["type"] = LIGHT_DIRECTIONAL
    l["doubleSide"] = 1
    return l
