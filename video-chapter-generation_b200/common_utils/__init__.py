from .set_random_seed import use_fix_random_seed  # noqa: F401
