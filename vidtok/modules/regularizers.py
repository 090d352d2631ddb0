from vidtok_b200.engine import DiagonalGaussianRegularizer, FSQRegularizer  # noqa: F401
