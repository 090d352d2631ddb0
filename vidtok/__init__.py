"""Import-path compatibility package: the reference's YAML configs name their classes by dotted path
(e.g. `target: vidtok.models.autoencoder.AutoencodingEngine`, configs/vidtok_kl_causal_488_4chn.yaml:3).
These thin modules re-export the vidtok_b200 implementations under the same paths so the reference's configs and
scripts resolve to the B200-native path unchanged."""
