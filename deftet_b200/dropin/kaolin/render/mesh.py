from deftet_b200.render import deftet_sparse_render  # noqa: F401
