from vidtok import overlay as _overlay

_overlay(__path__, "models")
