"""Model configuration: the reference's VelocityASRConfig (velocity_asr/model.py:23-68) and the
YAML -> config mapping that the reference keeps only inside scripts/train.py:158-174."""
from dataclasses import dataclass, fields
from typing import Any, Dict

SCAN_MODES = ("sequential", "parallel", "mamba")


@dataclass
class VelocityASRConfig:
    """Field names, order and defaults follow velocity_asr/model.py:27-63."""

    mel_bins: int = 80
    d_model: int = 192
    ssm_layers: int = 8
    ssm_state_dim: int = 64
    ssm_expand_ratio: int = 2
    ssm_kernel_size: int = 4
    global_ssm_layers: int = 2
    global_ssm_state_dim: int = 32
    attention_heads: int = 4
    attention_dim: int = 48
    vocab_size: int = 1000
    dropout: float = 0.1                  # inference only: identity
    gradient_checkpointing: bool = False  # training only: ignored
    scan_mode: str = "parallel"           # reference default (model.py:60); see DESIGN.md
    use_compile: bool = False             # no tracing compiler here: ignored

    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]) -> "VelocityASRConfig":
        """Unknown keys are dropped, as in model.py:65-68."""
        known = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in config_dict.items() if k in known})

    def to_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in fields(self)}


def config_from_yaml(path: str) -> VelocityASRConfig:
    """configs/model.yaml -> VelocityASRConfig, section by section as scripts/train.py:158-174
    reads it (keys the reference never reads, e.g. input.n_fft, stay unread)."""
    import yaml

    with open(path, "r") as fh:
        y = yaml.safe_load(fh) or {}

    def sec(name):
        return y.get(name, {}) or {}

    return VelocityASRConfig(
        mel_bins=sec("input").get("mel_bins", 80),
        d_model=sec("model").get("d_model", 192),
        ssm_layers=sec("ssm").get("num_layers", 8),
        ssm_state_dim=sec("ssm").get("state_dim", 64),
        ssm_expand_ratio=sec("ssm").get("expand_ratio", 2),
        ssm_kernel_size=sec("ssm").get("kernel_size", 4),
        global_ssm_layers=sec("global_context").get("ssm_layers", 2),
        global_ssm_state_dim=sec("global_context").get("ssm_state_dim", 32),
        attention_heads=sec("global_context").get("attention_heads", 4),
        attention_dim=sec("global_context").get("attention_dim", 48),
        vocab_size=sec("model").get("vocab_size", 1000),
        dropout=sec("model").get("dropout", 0.1),
        gradient_checkpointing=sec("memory").get("gradient_checkpointing", False),
        scan_mode=sec("performance").get("scan_mode", "parallel"),
        use_compile=sec("performance").get("use_compile", False),
    )
