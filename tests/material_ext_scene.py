"""A scene for the material-breadth extension (include/gdpt_wire.h): meshes of five and six surfaces (the
reference's BLASInstance holds three material ids, bvh.h:71), a roughness texture, a metallic texture and an albedo texture
flagged as sRGB (upstream samples every layer as UNORM, path_tracing_camera.cpp:182)."""
import numpy as np

from gdpathtracing_b200 import scenes


def _noise_layer(seed, size=64, lo=0, hi=255):
    r = scenes.splitmix64_floats(seed, size * size * 3).reshape(size, size, 3)
    img = np.empty((size, size, 4), np.uint8)
    img[..., :3] = (lo + r * (hi - lo)).astype(np.uint8)
    img[..., 3] = 255
    return img


def _faces_as_surfaces(half):
    """The five faces of an open box, one surface each."""
    hx, hy, hz = half
    quads = [
        scenes._quad_surface([[-hx, hy, -hz], [hx, hy, -hz], [hx, hy, hz], [-hx, hy, hz]], [0, 1, 0]),
        scenes._quad_surface([[hx, -hy, -hz], [hx, hy, -hz], [hx, hy, hz], [hx, -hy, hz]], [1, 0, 0]),
        scenes._quad_surface([[-hx, -hy, -hz], [-hx, hy, -hz], [-hx, hy, hz], [-hx, -hy, hz]], [-1, 0, 0]),
        scenes._quad_surface([[-hx, -hy, hz], [hx, -hy, hz], [hx, hy, hz], [-hx, hy, hz]], [0, 0, 1]),
        scenes._quad_surface([[-hx, -hy, -hz], [hx, -hy, -hz], [hx, hy, -hz], [-hx, hy, -hz]], [0, 0, -1]),
    ]
    return [scenes._merge([q]) for q in quads]


def material_ext_scene(ext=True, many_surfaces=True):
    """many_surfaces=False keeps every mesh at three surfaces or fewer, which is all the reference can address: the form
    used to compare the extension switched off with the reference itself."""
    sc = scenes.SceneDesc("material_ext", camera_transform12=scenes.transform12(None, (0.2, 0.3, 8.5)), fov=55.0, texture_array_resolution=64)
    sc.material_ext = ext
    sc.textures = [_noise_layer(31), _noise_layer(32, lo=20, hi=250), _noise_layer(33), scenes._checker(64, 8, 5)]
    sc.materials = [
        dict(albedo=(0.8, 0.8, 0.8), roughness=0.9),
        dict(albedo=(1.0, 1.0, 1.0), roughness=0.8, albedo_texture=0, albedo_srgb=True),                  # sRGB colour layer
        dict(albedo=(0.9, 0.6, 0.3), roughness=1.0, metallic=0.2, roughness_texture=1),                   # roughness map
        dict(albedo=(0.9, 0.9, 0.9), roughness=0.3, metallic=1.0, metallic_texture=2),                    # metallic map
        dict(albedo=(0.7, 0.9, 0.7), roughness=0.9, metallic=0.8, albedo_texture=3, roughness_texture=1, metallic_texture=2, albedo_srgb=True),
        dict(albedo=(1, 1, 1), emission=(1.0, 0.9, 0.8), emission_energy_multiplier=5.0),
        dict(albedo=(0.3, 0.4, 0.9), roughness=0.4, albedo_texture=3),                                    # UNORM colour layer
    ]
    sc.default_material = 0
    room = scenes._cornell_room()                                   # three surfaces
    box5 = _faces_as_surfaces((0.9, 0.9, 0.9))                      # five surfaces
    box6 = box5 + [scenes._merge([scenes._quad_surface([[-0.9, -0.9, -0.9], [0.9, -0.9, -0.9], [0.9, -0.9, 0.9], [-0.9, -0.9, 0.9]], [0, -1, 0])])]
    if not many_surfaces:
        box5, box6 = box5[:3], box6[3:]
    sc.meshes = [room, box5, box6]
    sc.instances = [
        dict(mesh=0, transform12=scenes.ROOM_TRANSFORM, surface_overrides=[1, 2, 3]),
        dict(mesh=1, transform12=scenes.transform12(None, (-1.3, -2.0, 0.2)), surface_overrides=[4, 3, 2, 1, 6]),
        dict(mesh=2, transform12=scenes.transform12([[0.8, 0.0, 0.6], [0.0, 1.0, 0.0], [-0.6, 0.0, 0.8]], (1.4, -1.2, -0.6)),
             surface_overrides=[2, 4, 6, 1, 3, 5]),
        dict(mesh=1, transform12=scenes.LIGHT_TRANSFORM, surface_overrides=[5, 5, 5, 5, 5]),
    ]
    return sc
