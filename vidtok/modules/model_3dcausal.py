from vidtok_b200.engine import DecoderCausal3DPadding, EncoderCausal3DPadding  # noqa: F401
