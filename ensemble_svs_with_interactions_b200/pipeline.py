"""Batched (song, track) synthesis: mgc + bap diffusion followed by the uSFGAN vocoder, many tracks per launch.

Replaces the reference's driver loop for this path (first part): ``nnsvs/bin/synthesis_multitrack.py
:113-118`` walks all utterance pairs and runs every model with batch size 1 (``multistream.py:1684-1696`` calls the mgc
and bap ``GaussianDiffusion.inference`` once per track, ``gen.py:1694`` the vocoder once per track).  Every (song, track)
item is independent once its conditioning exists, so here they are

* partitioned over the ranks with ``sharding.assign`` (no collective),
* grouped into padded batches under a frame budget with ``sharding.batches``,
* run through ``GaussianDiffusion.inference`` (all tracks of a batch per denoiser launch) and
  ``USFGANWrapper.inference_batch`` (all tracks per vocoder launch),
* trimmed back to their own lengths and returned in input order.

Between the models, on the device (second part): the GV post-filter of the mgc stream and the
zero-phase low-pass of both streams (``postprocess.variance_scaling`` / ``postprocess.lowpass_filter`` — gen.py:1394-1418,
1500-1513 do these per utterance and per dimension in numpy / scipy on the host).  What stays outside: the pyworld bap
round trip (gen.py:1639-1670), the learned post-filters and the feature scalers — they enter through ``aux_fn``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch

from . import postprocess, sharding


@dataclass
class BatchPlan:
    """One padded batch: ``items`` (indices into the caller's list), all padded to ``frames`` frames."""
    items: List[int]
    frames: int


def plan_batches(lengths: Sequence[int], max_frames: int, world_size: int = 1, rank: int = 0, multiple: int = 1) -> List[BatchPlan]:
    """This rank's batches.  Items go to ranks longest-first (``sharding.assign``); within the rank they are grouped so
    that ``len(batch) * padded_frames <= max_frames`` (a single item longer than the budget still gets its own batch).
    ``multiple``: pad the batch length up to a multiple of it (the reference pads to the models' reduction factor,
    util.py:77,163-164)."""
    if max_frames < 1 or multiple < 1:
        raise ValueError("max_frames and multiple must be positive")
    mine = sharding.assign(lengths, world_size)[rank]
    padded = [-(-int(n) // multiple) * multiple for n in lengths]
    return [BatchPlan(items=list(b), frames=max(padded[i] for i in b)) for b in sharding.batches(mine, padded, max_frames)]


def pair_items(utt_ids: Sequence[str]) -> List[tuple]:
    """The (main, interacting) utterance pairs the reference synthesises, in its order, without its O(N^2) scan.

    ``synthesis_multitrack.py:113-118`` walks every ordered pair of utterance ids ``<singer>_<segment...>`` and keeps those
    whose segment part (everything after the first ``_``) is equal — a track paired with every track of the same
    segment, itself included.  Here the ids are bucketed by segment once and the pairs are emitted bucket-wise in the
    same order (outer id in list order, inner ids in list order)."""
    buckets = {}
    for u in utt_ids:
        buckets.setdefault(tuple(u.split("_")[1:]), []).append(u)
    return [(u0, u1) for u0 in utt_ids for u1 in buckets[tuple(u0.split("_")[1:])]]


def _pad_time(x: torch.Tensor, frames: int, mode: str) -> torch.Tensor:
    """x [T, C] -> [frames, C]; ``replicate`` repeats the last frame (the reference's pad_inference), ``zeros`` appends 0."""
    T = x.shape[0]
    if T == frames:
        return x
    if T > frames:
        raise ValueError(f"item of {T} frames in a batch of {frames}")
    tail = x[-1:].expand(frames - T, -1) if mode == "replicate" else x.new_zeros((frames - T, x.shape[1]))
    return torch.cat([x, tail], dim=0)


class EnsembleSynthesizer:
    """mgc / bap diffusion + vocoder over batches of tracks.

    mgc, bap: ``GaussianDiffusion`` drop-ins (already on the device, eval mode), with or without an ``encoder``
    (``model.FFConvLSTM``: then ``cond_*`` are the linguistic features and every item's own length drives its packed
    BiLSTM); vocoder: ``USFGANWrapper``.
    aux_fn(mgc [B,T,M1], bap [B,T,M2], f0 [B,T,1]) -> vocoder aux features [B,T,C]; default: concatenate mgc and bap.
    max_frames: frame budget of one batch (tracks x padded frames).
    smoothing_cutoff: Hz, or None — ``trajectory_smoothing`` of gen.postprocess_acoustic (its default: 50 at
    ``frame_rate`` = 200 frames per second) applied to both streams.
    gv_mgc: global variance of the mgc stream [M1] (the scaler's ``var_``), or None — the GV post-filter, applied to
    the frames marked in ``note_masks`` (all valid frames when no masks are given), dimensions >= ``gv_offset``.
    out_scaler_mgc / out_scaler_bap: ``postprocess.StandardScaler`` / ``MinMaxScaler`` whose ``inverse_transform`` takes
    the streams back to feature units right after sampling (gen.py:1145); vocoder_in_scaler: its ``transform`` is applied
    to the assembled aux features (gen.predict_waveform).  All on the device.
    vuv: the V/UV stream's model (``model.FFConvLSTM`` in the recipe), or None.  As in the composite acoustic model
    (acoustic_models/multistream.py:1684-1720, main track) it reads ``cat([x, mgc, lf0])`` where ``cond_mgc = cat([x, lf0])``
    and ``mgc`` is the diffusion output before any scaler; frames whose (inverse-scaled) output is below
    ``vuv_threshold`` get f0 = 0 before the vocoder (gen.py: ``f0[vuv < vuv_threshold] = 0``)."""

    def __init__(self, mgc, bap, vocoder, max_frames: int = 36000,
                 aux_fn: Optional[Callable[[torch.Tensor, torch.Tensor, torch.Tensor], torch.Tensor]] = None,
                 smoothing_cutoff: Optional[float] = None, frame_rate: int = 200, gv_mgc: Optional[torch.Tensor] = None,
                 gv_offset: int = 2, out_scaler_mgc=None, out_scaler_bap=None, vocoder_in_scaler=None,
                 vuv=None, vuv_threshold: float = 0.5, out_scaler_vuv=None):
        self.mgc, self.bap, self.vocoder = mgc, bap, vocoder
        self.max_frames = int(max_frames)
        self.smoothing_cutoff, self.frame_rate = smoothing_cutoff, int(frame_rate)
        self.gv_mgc, self.gv_offset = gv_mgc, int(gv_offset)
        self.out_scaler_mgc, self.out_scaler_bap, self.vocoder_in_scaler = out_scaler_mgc, out_scaler_bap, vocoder_in_scaler
        self.vuv, self.vuv_threshold, self.out_scaler_vuv = vuv, float(vuv_threshold), out_scaler_vuv
        self.aux_fn = aux_fn if aux_fn is not None else (lambda m, b, f0: torch.cat([m, b], dim=-1))
        self._side = None

    def _encode(self, cm, cb, lens):
        """The two streams' encoders are independent and latency-bound (a BiLSTM recurrence on a few dozen SMs each):
        run the bap one on a side stream while the mgc one runs on the current stream."""
        em, eb = getattr(self.mgc, "encoder", None), getattr(self.bap, "encoder", None)
        if em is None or eb is None:
            return (cm if em is None else em(cm, lens)), (cb if eb is None else eb(cb, lens))
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=cm.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            hb = eb(cb, lens)
        hm = em(cm, lens)
        cur.wait_stream(self._side)
        hb.record_stream(cur)
        return hm, hb

    @torch.no_grad()
    def synthesize(self, cond_mgc: Sequence[torch.Tensor], cond_bap: Sequence[torch.Tensor], f0: Sequence[torch.Tensor],
                   world_size: int = 1, rank: int = 0,
                   note_masks: Optional[Sequence[torch.Tensor]] = None,
                   noise: Optional[dict] = None) -> List[Optional[torch.Tensor]]:
        """Per item i: cond_mgc[i] [T_i, H1], cond_bap[i] [T_i, H2], f0[i] [T_i, 1] (Hz, 0 = unvoiced), any device;
        note_masks[i] [T_i] bool (frames inside notes, for the GV post-filter).
        Returns the waveforms [T_i * hop] of this rank's items in input order (None for items of other ranks).
        ``noise`` (parity harness only; the call must form a single batch): {"mgc": (x_T, z), "bap": (x_T, z),
        "vocoder": {"sine": ..., "noise": ...}} replaces the Gaussian draws of the three models, tracks in plan order."""
        n = len(cond_mgc)
        if not (len(cond_bap) == n and len(f0) == n):
            raise ValueError("cond_mgc, cond_bap and f0 must have one entry per item")
        lengths = [int(c.shape[0]) for c in cond_mgc]
        for i in range(n):
            if cond_bap[i].shape[0] != lengths[i] or f0[i].shape[0] != lengths[i]:
                raise ValueError(f"item {i}: conditioning and f0 lengths differ")
        dev = next(self.mgc.parameters()).device
        hop = int(self.vocoder.config.data.hop_size)
        out: List[Optional[torch.Tensor]] = [None] * n
        plans = plan_batches(lengths, self.max_frames, world_size, rank)
        if noise is not None and len(plans) != 1:
            raise ValueError("noise injection needs all items in one batch")
        nz = noise or {}
        for plan in plans:
            cm = torch.stack([_pad_time(cond_mgc[i].to(dev, torch.float32), plan.frames, "replicate") for i in plan.items])
            cb = torch.stack([_pad_time(cond_bap[i].to(dev, torch.float32), plan.frames, "replicate") for i in plan.items])
            f = torch.stack([_pad_time(f0[i].to(dev, torch.float32), plan.frames, "zeros") for i in plan.items])
            lens = [lengths[i] for i in plan.items]         # used by the streams' encoders (packed BiLSTM), if any
            hm, hb = self._encode(cm, cb, lens)
            xm, zm = nz.get("mgc", (None, None))
            xb_, zb_ = nz.get("bap", (None, None))
            m = self.mgc.inference(hm, cond_is_encoded=True, x_T=xm, z=zm)   # [B, T, M1]
            b = self.bap.inference(hb, cond_is_encoded=True, x_T=xb_, z=zb_)   # [B, T, M2]
            if self.vuv is not None:
                v = self.vuv.inference(torch.cat([cm[..., :-1], m, cm[..., -1:]], dim=-1).contiguous(), lens)   # [B, T, 1]
                if self.out_scaler_vuv is not None:
                    v = self.out_scaler_vuv.inverse_transform(v)
                f = torch.where(v < self.vuv_threshold, torch.zeros_like(f), f)
            if self.out_scaler_mgc is not None:
                m = self.out_scaler_mgc.inverse_transform(m)
            if self.out_scaler_bap is not None:
                b = self.out_scaler_bap.inverse_transform(b)
            if self.gv_mgc is not None:                        # gen.py:1394-1418
                mask = None
                if note_masks is not None:
                    mask = torch.stack([_pad_time(note_masks[i].to(dev).to(torch.uint8)[:, None], plan.frames, "zeros")[:, 0]
                                        for i in plan.items])
                m = postprocess.variance_scaling(self.gv_mgc, m, offset=self.gv_offset, note_mask=mask, lengths=lens)
            if self.smoothing_cutoff is not None:              # gen.py:1500-1513
                M1 = m.shape[-1]                               # one launch for both streams: every trajectory is independent
                mb = postprocess.lowpass_filter(torch.cat([m, b], dim=-1), self.frame_rate, cutoff=self.smoothing_cutoff, lengths=lens)
                m, b = mb[..., :M1], mb[..., M1:]
            aux = self.aux_fn(m, b, f).contiguous()
            if self.vocoder_in_scaler is not None:
                aux = self.vocoder_in_scaler.transform(aux)
            if "vocoder" in nz:
                wav = self.vocoder.inference_batch(f, aux, noise=nz["vocoder"])
            else:
                wav = self.vocoder.inference_batch(f, aux)       # [B, 1, T * hop]
            for k, i in enumerate(plan.items):
                out[i] = wav[k, 0, :lengths[i] * hop].clone()
        return out
