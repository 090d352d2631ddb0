from vidtok_b200.compat_util import compute_psnr, get_obj_from_str, instantiate_from_config, print0  # noqa: F401
