"""Parameter container with the reference's module tree, state_dict keys and random init.

The reference builds nn.Modules whose forward() does the arithmetic; here the modules only
*hold* parameters (the arithmetic runs in libvasr.so), but they are created in the same order
and re-initialised by the same rule as VELOCITYASR._init_weights (model.py:305-318), so that
`torch.manual_seed(s); VELOCITYASR(cfg)` draws bit-identical weights and `state_dict()` has the
same 208 keys in the same order.  That is what lets parity tests on a box
without the reference rebuild the reference's random-init weights from a seed.
"""
import math

import torch
import torch.nn as nn


class ParamGroup(nn.Module):
    """A named bag of parameters / sub-groups.  Not callable: compute lives in the CUDA library."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container only; call the VELOCITYASR model")


def _indexed(mods: dict) -> nn.ModuleDict:
    """Sub-modules under numeric names, e.g. ffn.0 / ffn.3 (nn.Sequential slots with parameters)."""
    return nn.ModuleDict({str(i): m for i, m in mods.items()})


def time_table(rows: int, d_model: int) -> torch.Tensor:
    """Sinusoidal rows of pe_time (model.py:94-100); any number of rows (the reference stops at 5000)."""
    half = d_model // 2
    pe = torch.zeros(rows, half)
    pos = torch.arange(0, rows, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, half, 2).float() * (-math.log(10000.0) / half))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def _ssm_block(d_model, state_dim, expand, ksize) -> ParamGroup:
    di = d_model * expand
    blk = ParamGroup()
    blk.norm1 = nn.LayerNorm(d_model)
    blk.norm2 = nn.LayerNorm(d_model)
    blk.conv = nn.Conv1d(d_model, d_model, kernel_size=ksize, padding=ksize - 1, groups=d_model)
    ssm = ParamGroup()
    ssm.in_proj = nn.Linear(d_model, 2 * di, bias=False)
    ssm.x_proj = nn.Linear(di, 2 * state_dim, bias=False)
    ssm.dt_proj = nn.Linear(di, di, bias=True)
    ssm.A_log = nn.Parameter(torch.log(torch.arange(1, state_dim + 1, dtype=torch.float32)))
    ssm.D = nn.Parameter(torch.ones(di))
    ssm.out_proj = nn.Linear(di, d_model, bias=False)
    blk.ssm = ssm
    blk.ffn = _indexed({0: nn.Linear(d_model, di), 3: nn.Linear(di, d_model)})
    return blk


def _ssm_stack(d_model, layers, state_dim, expand, ksize) -> ParamGroup:
    st = ParamGroup()
    st.layers = nn.ModuleList([_ssm_block(d_model, state_dim, expand, ksize) for _ in range(layers)])
    st.norm = nn.LayerNorm(d_model)
    return st


def build_parameter_tree(model: nn.Module, cfg) -> None:
    """Attach temporal_binding / local_ssm / global_context / ctc_head to `model`."""
    d = cfg.d_model
    tb = ParamGroup()
    tb.conv = nn.Conv1d(cfg.mel_bins, d, kernel_size=3, stride=2, padding=1)
    pos = ParamGroup()
    pos.register_buffer("pe_time", time_table(5000, d))
    pos.pe_freq = nn.Parameter(torch.randn(1, 1, d // 2) * 0.02)
    tb.pos_encoding = pos
    tb.norm = nn.LayerNorm(d)
    model.temporal_binding = tb

    model.local_ssm = _ssm_stack(d, cfg.ssm_layers, cfg.ssm_state_dim, cfg.ssm_expand_ratio, cfg.ssm_kernel_size)

    gc = ParamGroup()
    p1 = ParamGroup()
    p1.pool_proj = nn.Linear(d, d)
    gc.pool1 = p1
    # GlobalSSM hard-codes expand_ratio=2, kernel_size=4 (ssm.py:529-538)
    gc.global_ssm = _ssm_stack(d, cfg.global_ssm_layers, cfg.global_ssm_state_dim, 2, 4)
    p2 = ParamGroup()
    p2.pool_proj = nn.Linear(d, d)
    gc.pool2 = p2
    att = ParamGroup()
    att.q_proj = nn.Linear(d, cfg.attention_dim)
    att.k_proj = nn.Linear(d, cfg.attention_dim)
    att.v_proj = nn.Linear(d, cfg.attention_dim)
    att.out_proj = nn.Linear(cfg.attention_dim, d)
    gc.cross_attention = att
    gc.norm1 = nn.LayerNorm(d)
    gc.norm2 = nn.LayerNorm(d)
    fu = ParamGroup()
    fu.gate_proj = _indexed({0: nn.Linear(2 * d, d)})
    fu.local_proj = nn.Linear(d, d)
    fu.global_proj = nn.Linear(d, d)
    fu.out_proj = nn.Linear(d, d)
    gc.fusion = fu
    model.global_context = gc

    head = ParamGroup()
    head.proj = _indexed({0: nn.LayerNorm(d), 2: nn.Linear(d, cfg.vocab_size)})
    model.ctc_head = head


def reference_init_(model: nn.Module) -> None:
    """The rule of VELOCITYASR._init_weights (model.py:305-318), applied in module order."""
    for m in model.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv1d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
