from vidtok_b200.engine import AutoencodingEngine  # noqa: F401  (v1.0 engine)
