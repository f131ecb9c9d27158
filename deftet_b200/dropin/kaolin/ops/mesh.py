from deftet_b200.render import check_sign  # noqa: F401
