"""CPU restatement of the global CNN feature extractor -- TEST INFRASTRUCTURE ONLY (imported by tests/ and the
fixture scripts; the product path never touches it).

`GlobalFeatureExtractorCNN.forward` of /root/reference/src/feature_extractors.py:24-34, op for op in plain torch:
normalise by the batch-wide max |u| (:26), `selu(conv(u))` per layer (:27-28; Conv1d / Conv2d, kernel 3, stride 1,
padding 1, built at :16-21), adaptive average pool to one value per channel (:29-30), flatten (:31).
`reshape_fd_tensor_to_grid` (src/utils_data.py:125-141) is restated next to it because the CUDA path fuses it into
the kernel's load.  Pinned by tests/golden_cnn/*.pt, minted from the reference's own class
(oracle/ref_harness/make_golden_cnn.py)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def cnn_features(u: torch.Tensor, weights, biases) -> torch.Tensor:
    """u [B, 1, H, W] or [B, 1, W]; weights / biases: per layer, torch Conv layout."""
    conv = F.conv1d if u.dim() == 3 else F.conv2d
    u = u / torch.max(torch.abs(u))
    for w, b in zip(weights, biases):
        u = F.selu(conv(u, w, b, stride=1, padding=1))
    return u.mean(dim=tuple(range(2, u.dim()))).view(u.size(0), -1)


def reshape_fd_tensor_to_grid(u_true, mapping_tensor, mesh_dims, batch_size=1, dim=None):
    """src/utils_data.py:125-141."""
    if dim == 1:
        return u_true.reshape(batch_size, -1)
    g = u_true.reshape(batch_size, -1)
    g = torch.gather(g, 1, mapping_tensor.unsqueeze(0).expand(batch_size, -1))
    g = g.reshape(batch_size, mesh_dims[0], mesh_dims[1])
    return torch.flip(torch.transpose(g, 1, 2), [1])


def grid_gather_index(mapping_tensor: torch.Tensor, n: int) -> torch.Tensor:
    """The composite of the three reorderings above as ONE index map: grid[b, r, c] = u[b, index[r * n + c]]."""
    probe = torch.arange(n * n, dtype=torch.float64).view(1, -1)
    return reshape_fd_tensor_to_grid(probe, mapping_tensor, [n, n], 1, 2).reshape(-1).long()
