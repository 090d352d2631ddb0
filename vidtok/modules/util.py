from vidtok_b200.compat_util import (compute_psnr, compute_ssim, default, exists, get_obj_from_str, get_valid_dirs,  # noqa: F401
                                     get_valid_paths, instantiate_from_config, isheatmap, print0, seed_anything)
