from vidtok_b200.engine import AutoencodingEngineV11 as AutoencodingEngine  # noqa: F401
