"""Test stub: imageio is not installed in this image; diff_render 2_data/load_blender.py imports it at module level and the
optimisation script writes a preview video with it at the very end."""


def imread(*a, **k):
    raise RuntimeError("imageio stub: no image files in the synthetic test")


def mimwrite(*a, **k):
    return None
