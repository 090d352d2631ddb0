from vidtok_b200.engine import Decoder3D, Encoder3D  # noqa: F401
