from vidtok_b200.engine import DecoderCausal3DPaddingV11 as DecoderCausal3DPadding  # noqa: F401
from vidtok_b200.engine import EncoderCausal3DPaddingV11 as EncoderCausal3DPadding  # noqa: F401
