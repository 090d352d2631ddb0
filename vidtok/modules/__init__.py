from vidtok import overlay as _overlay

_overlay(__path__, "modules")   # vidtok.modules.lpips, .logger, ... resolve to the reference when it is importable
