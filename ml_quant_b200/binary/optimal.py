"""Optimal v1 of the 2-bit / ternary least-squares quantizer.

Mirror of quant/binary/optimal.py.  ``opt_v1`` (:121-155) is one CUDA kernel launch (csrc/lsq_solve.cu,
sort-free, no host synchronisation).  ``compute_mask`` (:41-83) and ``cost_function`` (:16-38) are the
reference's inspection helpers; they are kept as device-side torch expressions for API completeness and
are not used by any forward pass here.

Prefix sums: the reference calls ``values.cumsum(dim=1)`` on an fp32 tensor; on the CPU, where its results are
pinned, ATen accumulates that in DOUBLE and rounds each prefix to fp32 (tests/test_oracle.py::
test_cpu_cumsum_accumulates_in_double).  ``compute_mask`` therefore forms the prefix sums in fp64 and rounds them to
fp32 -- the reference's CPU arithmetic, not a refinement of it -- and reproduces the golden candidate sets bit for
bit on the GPU (tests/test_gpu_round2.py::test_compute_mask_and_cost_function_against_golden).
"""
from typing import Tuple

import torch

from .. import ops
from .ste import binary_sign


def opt_v1(matrix: torch.Tensor, ternary: bool, skip: int = 1) -> torch.Tensor:
    """v1 per row of a 2-D tensor, shape [rows, 1]; only every ``skip``-th column enters the solve."""
    ops.require_cuda(matrix, 'matrix')
    with torch.no_grad():
        m2 = matrix.reshape(matrix.shape[0], -1)
        return ops.solve_v1(m2, ternary, skip).view(-1, 1)


def cost_function(matrix: torch.Tensor, v1s: torch.Tensor, ternary: bool = False) -> torch.Tensor:
    """||r - v2 sign(r)||_2 with r = |m| - v1 for every candidate column of ``v1s`` -> [rows, cands]."""
    ops.require_cuda(matrix, 'matrix')
    rows = matrix.shape[0]
    out = []
    for j in range(v1s.shape[1]):           # one candidate at a time: never materialise [rows, cands, n]
        c = v1s[:, j].reshape(rows, 1)
        r = matrix - c * binary_sign(matrix)
        v2 = c if ternary else r.abs().mean(dim=-1, keepdim=True)
        out.append(torch.norm(r - v2 * binary_sign(r), dim=-1))
    return torch.stack(out, dim=1)


def compute_mask(matrix: torch.Tensor, ternary: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Candidate mask over the interior sorted positions and the selected values (fp32, device)."""
    ops.require_cuda(matrix, 'matrix')
    srt = torch.sort(matrix, dim=1).values
    n = matrix.shape[1]
    pre = srt.double().cumsum(dim=1).float()
    k = torch.arange(1, n + 1, device=matrix.device)
    rest = (n - k).clamp(min=1)
    half = 0.5 * ((pre[:, -1:] - pre) / rest)[:, 1:-1]
    lo, nx = srt[:, 1:-1], srt[:, 2:]
    mask = (lo <= half) & (half <= nx)
    if not ternary:
        mid = 0.5 * ((pre / k)[:, 1:-1] + ((pre[:, -1:] - pre) / rest)[:, 1:-1])
        mask = mask | ((lo <= mid) & (mid <= nx))
    return mask, torch.masked_select(lo, mask)
