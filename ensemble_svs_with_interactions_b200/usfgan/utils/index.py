"""Indexing helpers — drop-in for ``nnsvs.usfgan.utils.index`` (nnsvs/usfgan/utils/index.py:12-84)."""
import torch

from ... import ops


def pd_indexing(x, d, dilation, batch_index=None, ch_index=None):
    """Pitch-dependent past/future taps (index.py:12-54) -> (xP, xF), each (B, C, T).

    The reference pads two copies of ``x`` and does two advanced-index gathers; here the tap indices come from
    svsk_pd_index (same fp32 round-half-even arithmetic) and the gather is an identity-weight INDEXED conv tap.
    ``batch_index`` / ``ch_index`` are accepted for signature compatibility and ignored.
    """
    idx = ops.pd_index(d.to(torch.float32).contiguous(), dilation)
    C = x.shape[1]
    eye = torch.eye(C, device=x.device, dtype=torch.float32)
    zero = torch.zeros_like(eye)
    wP = torch.stack([eye, zero, zero], dim=2).contiguous()
    wF = torch.stack([zero, zero, eye], dim=2).contiguous()
    x = x.to(torch.float32).contiguous()
    return (ops.conv1d_f32(x, wP, pad_mode=ops.PAD_INDEXED, idx=idx),
            ops.conv1d_f32(x, wF, pad_mode=ops.PAD_INDEXED, idx=idx))


def index_initial(n_batch, n_ch, tensor=True):
    """index.py:57-84 (kept for API parity; the kernels do not need explicit batch/channel index tensors)."""
    batch_index = [[[i]] * n_ch for i in range(n_batch)]
    ch_index = [[[i] for i in range(n_ch)]] * n_batch
    if tensor:
        batch_index = torch.tensor(batch_index)
        ch_index = torch.tensor(ch_index)
        if torch.cuda.is_available():
            batch_index = batch_index.cuda()
            ch_index = ch_index.cuda()
    return batch_index, ch_index
