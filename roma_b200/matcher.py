"""`RegressionMatcher`: the reference's public matcher API on top of the B200 engine.

Mirrors `romatch.models.matcher.RegressionMatcher` (`romatch/models/matcher.py:550-986`): same
constructor-level attributes (mutable, as the reference's README documents), same method names,
argument meaning, return shapes/dtypes and error behaviour, so that callers (`demo/*.py`,
`romatch/benchmarks/*`) can switch implementation without edits.  The heavy lifting (`match`, `forward`,
the KDE inside `sample`) runs in the hand-written CUDA kernels behind `roma_b200.engine.Engine`; the small
geometry helpers are plain tensor arithmetic on whatever device their inputs live on.
"""
from __future__ import annotations

import math
import os
from warnings import warn

import torch
import torch.nn.functional as F
from PIL import Image

from . import cabi
from .engine import Engine
from .preprocess import DevicePreprocessor, check_input


class RegressionMatcher:
    def __init__(self, engine: Engine, h=448, w=448, sample_mode="threshold_balanced", upsample_preds=False,
                 symmetric=False, sample_thresh=0.05, name=None, attenuate_cert=None, upsample_res=None):
        self.engine = engine
        self.attenuate_cert = attenuate_cert
        self.name = name
        self.w_resized = w
        self.h_resized = h
        self.sample_mode = sample_mode
        self.upsample_preds = upsample_preds
        self.upsample_res = upsample_res or (14 * 16 * 6, 14 * 16 * 6)      # matcher.py:575
        self.symmetric = symmetric
        self.sample_thresh = sample_thresh
        self.training = False
        self.device_sampler = True      # sample(): weighted sampling without replacement in one kernel per draw (False: torch.multinomial)
        self.use_cuda_graph = True      # replay the whole device side of match() as one CUDA graph per input shape
        self._graphs = {}
        self._pre = None
        self._sample_state = {}         # static buffers + CUDA graph of the device sampler, per (n, num, mode)
        self.graph_launches = 0         # kernels launched through graph replays (cabi.kernel_launches counts eager ones)

    # ---- nn.Module-ish conveniences callers rely on ------------------------------------------------
    def train(self, mode: bool = True):
        self.training = False      # inference-only implementation; match() forces eval (matcher.py:790)
        return self

    def eval(self):
        return self.train(False)

    def to(self, *args, **kwargs):
        return self

    def _get_device(self):
        return self.engine.device

    def free_buffers(self):
        """Release every cached activation buffer and the CUDA graphs recorded over them."""
        self._graphs.clear()
        self._sample_state.clear()
        self.engine.free_buffers()

    def get_output_resolution(self):
        if not self.upsample_preds:
            return self.h_resized, self.w_resized
        return self.upsample_res

    # ---- dense matching -----------------------------------------------------------------------------
    def _states_to_corresps(self, states):
        return {s: {"flow": st[..., :2].permute(0, 3, 1, 2), "certainty": st[..., 2:3].permute(0, 3, 1, 2)}
                for s, st in states.items()}

    @torch.inference_mode()
    def forward(self, batch, batched=True, upsample=False, scale_factor=1):
        """{scale: {"flow" [B,2,h,w], "certainty" [B,1,h,w]}} like `RegressionMatcher.forward` (matcher.py:631-652)."""
        return self._forward(batch, False, upsample, scale_factor)

    @torch.inference_mode()
    def forward_symmetric(self, batch, batched=True, upsample=False, scale_factor=1):
        return self._forward(batch, True, upsample, scale_factor)

    def _forward(self, batch, symmetric, upsample, scale_factor):
        eng = self.engine
        im_a = batch["im_A"].to(eng.device, torch.float32)
        im_b = batch["im_B"].to(eng.device, torch.float32)
        images = torch.cat((im_a, im_b)).contiguous()
        state_in = None
        if upsample:
            c = batch["corresps"]
            st = torch.cat((c["flow"], c["certainty"]), dim=1).permute(0, 2, 3, 1).contiguous().float()
            state_in = (st, st.shape[1], st.shape[2])
        with torch.cuda.device(eng.device):
            _, states, _ = eng.run_pass(images, im_a.shape[0], symmetric, upsample, float(scale_factor), state_in,
                                        keep_states=True)
        return self._states_to_corresps(states)

    def _preprocessor(self) -> DevicePreprocessor:
        if self._pre is None:
            self._pre = DevicePreprocessor(self.engine.device)
        return self._pre

    @torch.inference_mode()
    def match(self, im_A_input, im_B_input, *args, im_A_high_res=None, im_B_high_res=None, batched=True, device=None):
        """Dense warp and certainty (matcher.py:779-934).  Returns (warp [b,H,W*(2 if symmetric),4] fp32 in
        [-1,1], certainty [b,H,W*(2)] fp32 in [0,1]); extra positional args are ignored like the reference."""
        if not batched:
            raise ValueError("batched must be True, non-batched inference is no longer supported.")
        eng = self.engine
        if device is None:
            device = eng.device
        if torch.device(device).type != "cuda":
            raise RuntimeError("roma_b200 computes on CUDA only; device=%r" % (device,))
        im_A = check_input(im_A_input)
        im_B = check_input(im_B_input)
        symmetric = self.symmetric
        ws, hs = self.w_resized, self.h_resized
        scale_factor = math.sqrt(hs * ws / (560 ** 2))
        pil_route = isinstance(im_A, Image.Image) and isinstance(im_B, Image.Image)
        if pil_route:
            # raw RGB bytes go up once per image; Pillow's bicubic resize + normalisation run on the device (csrc/preprocess.cu)
            b = 1
            pre = self._preprocessor()
            raw_a, raw_b = pre.upload(im_A), pre.upload(im_B)
            a_t = pre.resize_normalize(raw_a, (hs, ws))[None]
            b_t = pre.resize_normalize(raw_b, (hs, ws))[None]
        elif isinstance(im_A, torch.Tensor) and isinstance(im_B, torch.Tensor):
            b, c, h, w = im_A.shape
            b, c, h2, w2 = im_B.shape
            assert w == w2 and h == h2, "For batched images we assume same size"
            if h != self.h_resized or self.w_resized != w:
                warn("Model resolution and batch resolution differ, may produce unexpected results")
            hs, ws = h, w
            a_t, b_t = im_A, im_B
        else:
            raise ValueError(f"Unsupported input type: {type(im_A)=} and {type(im_B)=}")

        a_h = b_h = None
        if self.upsample_preds:
            hs, ws = self.upsample_res
            if im_A_high_res is None and im_B_high_res is None:
                # the reference re-opens / re-uses the same two images here (matcher.py:855-866): same bytes, already on the device
                if not isinstance(im_A_input, (str, os.PathLike)):
                    assert isinstance(im_A_input, Image.Image), f"Unsupported input type: {type(im_A_input)=}"
                    assert isinstance(im_B_input, Image.Image), f"Unsupported input type: {type(im_B_input)=}"
                assert pil_route, "upsample_preds without high-res tensors needs path or PIL inputs"
                a_h = pre.resize_normalize(raw_a, (hs, ws))[None]
                b_h = pre.resize_normalize(raw_b, (hs, ws))[None]
            elif im_A_high_res is not None and im_B_high_res is not None:
                a_h, b_h = im_A_high_res, im_B_high_res
            else:
                raise ValueError(f"Invalid upsample_preds and high_res inputs with {im_A=},{im_A_high_res=},{im_B=} and {im_B_high_res=}")
        with torch.cuda.device(eng.device):
            return self._match_device(a_t, b_t, a_h, b_h, b, symmetric, scale_factor)

    def _run_device(self, images, images_hi, b, symmetric, scale_factor, attenuate, warp, cert):
        """Both passes + epilogue on the current stream; no allocation, no host sync (CUDA-graph capturable)."""
        sf_hi = math.sqrt(self.upsample_res[0] * self.upsample_res[1] / (560 ** 2))
        self.engine.run_match(images, images_hi, b, symmetric, scale_factor, sf_hi, attenuate, warp, cert)

    def _match_device(self, a_t, b_t, a_h, b_h, b, symmetric, scale_factor):
        eng = self.engine
        dev = eng.device
        hs, ws = a_t.shape[-2:]
        ho, wo = (a_h.shape[-2:] if a_h is not None else (hs, ws))
        wout = 2 * wo if symmetric else wo
        attenuate = bool(self.attenuate_cert)
        use_graph = self.use_cuda_graph and eng.debug is None and eng.profile is None and eng.gemm_profile is None
        key = (b, hs, ws, ho if a_h is not None else 0, wo if a_h is not None else 0, symmetric, attenuate, float(scale_factor),
               tuple(self.upsample_res))
        entry = self._graphs.get(key) if use_graph else None
        if entry is not None and entry["generation"] != eng.generation:
            # Engine.free_buffers() released the activation buffers this graph's kernels point into: re-record
            del self._graphs[key]
            entry = None
        if entry is None:
            images = torch.empty(2 * b, 3, hs, ws, dtype=torch.float32, device=dev)
            images_hi = torch.empty(2 * b, 3, ho, wo, dtype=torch.float32, device=dev) if a_h is not None else None
            warp = torch.empty(b, ho, wout, 4, dtype=torch.float32, device=dev)
            cert = torch.empty(b, ho, wout, dtype=torch.float32, device=dev)
            entry = dict(images=images, images_hi=images_hi, warp=warp, cert=cert, graph=None, calls=0, generation=eng.generation)
            if use_graph:
                self._graphs[key] = entry
        entry["images"][:b].copy_(a_t, non_blocking=True)
        entry["images"][b:].copy_(b_t, non_blocking=True)
        if a_h is not None:
            entry["images_hi"][:b].copy_(a_h, non_blocking=True)
            entry["images_hi"][b:].copy_(b_h, non_blocking=True)
        args = (entry["images"], entry["images_hi"], b, symmetric, scale_factor, attenuate, entry["warp"], entry["cert"])
        if not use_graph:
            self._run_device(*args)
            return entry["warp"], entry["cert"]
        entry["calls"] += 1
        if entry["graph"] is None:
            self._run_device(*args)                 # eager: also allocates every activation buffer
            if entry["calls"] >= 2:                 # second call with this shape: capture for all later calls
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                launches0 = cabi.kernel_launches()
                with torch.cuda.graph(graph):
                    self._run_device(*args)
                entry["graph"] = graph
                entry["launches"] = cabi.kernel_launches() - launches0
        else:
            entry["graph"].replay()
            self.graph_launches += entry["launches"]
        return entry["warp"].clone(), entry["cert"].clone()

    # ---- sampling (matcher.py:598-629) ----------------------------------------------------------------
    def sample(self, matches, certainty, num=10000):
        """Certainty-thresholded, density-balanced match sampling (matcher.py:598-629).  Both weighted draws without
        replacement run on the device (`romab200_weighted_sample`: exponential race + radix select, the certainty
        thresholding and the density balancing fused into the key computation) and the 4*num x 4*num Gaussian KDE in
        `romab200_kde_density` without materialising the matrix; the seeds of the two draws come from torch's CPU generator,
        so `torch.manual_seed` makes the result reproducible.  With `device_sampler = False` the two `torch.multinomial`
        calls of the reference are used instead (same distribution, torch's RNG stream)."""
        if self.device_sampler and matches.is_cuda:
            return self._sample_device(matches, certainty, num)
        if "threshold" in self.sample_mode:
            upper_thresh = self.sample_thresh
            certainty = certainty.clone()
            certainty[certainty > upper_thresh] = 1
        matches, certainty = matches.reshape(-1, 4), certainty.reshape(-1)
        expansion_factor = 4 if "balanced" in self.sample_mode else 1
        good_samples = torch.multinomial(certainty, num_samples=min(expansion_factor * num, len(certainty)), replacement=False)
        good_matches, good_certainty = matches[good_samples], certainty[good_samples]
        if "balanced" not in self.sample_mode:
            return good_matches, good_certainty
        if good_matches.device.type != "cuda":
            raise RuntimeError("roma_b200.sample needs CUDA tensors (no CPU fallback)")
        with torch.cuda.device(good_matches.device):
            density = self.engine.kde(good_matches, std=0.1, half=True).to(torch.float16)   # kde.py: x.half()
        p = 1 / (density + 1)
        p[density < 10] = 1e-7      # at least ~10 perfect neighbours, as in the reference
        balanced_samples = torch.multinomial(p, num_samples=min(num, len(good_certainty)), replacement=False)
        return good_matches[balanced_samples], good_certainty[balanced_samples]

    def _sample_device(self, matches, certainty, num):
        """Device sampler; from the second call with the same sizes on, the whole chain (two draws, sort, gathers, KDE) is one
        CUDA-graph replay fed through static buffers, with the two seeds of a call written to a device word."""
        balanced = "balanced" in self.sample_mode
        thresholded = "threshold" in self.sample_mode
        dev = matches.device
        with torch.cuda.device(dev):
            n = certainty.numel()
            key = (n, num, self.sample_mode, float(self.sample_thresh), dev.index)
            st = self._sample_state.get(key)
            if st is None:
                k1 = min((4 if balanced else 1) * num, n)
                st = dict(m=torch.empty(n, 4, device=dev), c=torch.empty(n, device=dev), seeds=torch.zeros(2, dtype=torch.int64, device=dev),
                          seeds_host=torch.zeros(2, dtype=torch.int64).pin_memory(), idx1=torch.empty(k1, dtype=torch.int32, device=dev),
                          idx2=torch.empty(min(num, k1), dtype=torch.int32, device=dev), keys=torch.empty(n, device=dev),
                          scratch=torch.empty(2056, dtype=torch.int32, device=dev), k1=k1, graph=None, calls=0, out=None)
                self._sample_state[key] = st
            st["m"].copy_(matches.reshape(-1, 4), non_blocking=True)
            st["c"].copy_(certainty.reshape(-1), non_blocking=True)
            st["seeds_host"].copy_(torch.randint(0, 2 ** 62, (2,), dtype=torch.int64))       # CPU generator: follows torch.manual_seed
            st["seeds"].copy_(st["seeds_host"], non_blocking=True)

            def chain():
                m, c, k1 = st["m"], st["c"], st["k1"]
                cabi.call("romab200_weighted_sample", "rb_sample_args", values=c, n=n, k=k1, batch=1, stride=n, seed=0, seed_dev=st["seeds"],
                          transform=cabi.SAMPLE_THRESHOLD if thresholded else cabi.SAMPLE_IDENTITY, param=float(self.sample_thresh),
                          out_idx=st["idx1"], out_weights=None, keys=st["keys"], scratch=st["scratch"])
                sel1 = st["idx1"].long().sort().values           # the compaction order is not deterministic; the drawn SET is
                good_matches = m[sel1]
                w1 = torch.where(c[sel1] > self.sample_thresh, torch.ones((), device=dev), c[sel1]) if thresholded else c[sel1]
                if not balanced:
                    return good_matches, w1
                density = self.engine.kde(good_matches, std=0.1, half=True).to(torch.float16).float().contiguous()     # kde.py: x.half()
                cabi.call("romab200_weighted_sample", "rb_sample_args", values=density, n=k1, k=st["idx2"].numel(), batch=1, stride=k1, seed=0,
                          seed_dev=st["seeds"][1:], transform=cabi.SAMPLE_BALANCE, param=0.0, out_idx=st["idx2"], out_weights=None, keys=st["keys"],
                          scratch=st["scratch"])
                sel = st["idx2"].long().sort().values
                return good_matches[sel], w1[sel]

            st["calls"] += 1
            if st["graph"] is not None:
                st["graph"].replay()
                return st["out"][0].clone(), st["out"][1].clone()
            out = chain()
            if self.use_cuda_graph and st["calls"] >= 2:
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    st["out"] = chain()
                st["graph"] = graph
            return out

    # ---- small geometry helpers (matcher.py:672-773) ---------------------------------------------------
    def _to_pixel_coordinates(self, coords, H, W):
        return torch.stack((W / 2 * (coords[..., 0] + 1), H / 2 * (coords[..., 1] + 1)), dim=-1)

    def to_pixel_coordinates(self, coords, H_A, W_A, H_B=None, W_B=None):
        if coords.shape[-1] == 2:
            return self._to_pixel_coordinates(coords, H_A, W_A)
        if isinstance(coords, (list, tuple)):
            kpts_A, kpts_B = coords[0], coords[1]
        else:
            kpts_A, kpts_B = coords[..., :2], coords[..., 2:]
        return self._to_pixel_coordinates(kpts_A, H_A, W_A), self._to_pixel_coordinates(kpts_B, H_B, W_B)

    def to_normalized_coordinates(self, coords, H_A, W_A, H_B, W_B):
        if isinstance(coords, (list, tuple)):
            kpts_A, kpts_B = coords[0], coords[1]
        else:
            kpts_A, kpts_B = coords[..., :2], coords[..., 2:]
        kpts_A = torch.stack((2 / W_A * kpts_A[..., 0] - 1, 2 / H_A * kpts_A[..., 1] - 1), dim=-1)
        kpts_B = torch.stack((2 / W_B * kpts_B[..., 0] - 1, 2 / H_B * kpts_B[..., 1] - 1), dim=-1)
        return kpts_A, kpts_B

    def conf_from_fb_consistency(self, flow_forward, flow_backward, th=2):
        has_batch = flow_forward.dim() != 3
        if not has_batch:
            flow_forward, flow_backward = flow_forward[None], flow_backward[None]
        H, W = flow_forward.shape[-3:-1]
        th_n = 2 * th / max(H, W)
        xs = torch.linspace(-1 + 1 / W, 1 - 1 / W, W)
        ys = torch.linspace(-1 + 1 / H, 1 - 1 / H, H)
        coords = torch.stack(torch.meshgrid(xs, ys, indexing="xy"), dim=-1).to(flow_forward.device)
        coords_fb = F.grid_sample(flow_backward.permute(0, 3, 1, 2), flow_forward, align_corners=False,
                                  mode="bilinear").permute(0, 2, 3, 1)
        in_th = ((coords - coords_fb).norm(dim=-1) < th_n).float()
        return in_th if has_batch else in_th[0]

    def match_keypoints(self, x_A, x_B, warp, certainty, return_tuple=True, return_inds=False, max_dist=0.005, cert_th=0):
        x_A_to_B = F.grid_sample(warp[..., -2:].permute(2, 0, 1)[None], x_A[None, None], align_corners=False,
                                 mode="bilinear")[0, :, 0].mT
        cert_A_to_B = F.grid_sample(certainty[None, None, ...], x_A[None, None], align_corners=False,
                                    mode="bilinear")[0, 0, 0]
        D = torch.cdist(x_A_to_B, x_B)
        mutual = (D == D.min(dim=-1, keepdim=True).values) * (D == D.min(dim=-2, keepdim=True).values)
        inds_A, inds_B = torch.nonzero(mutual * (cert_A_to_B[:, None] > cert_th) * (D < max_dist), as_tuple=True)
        if return_tuple:
            return (inds_A, inds_B) if return_inds else (x_A[inds_A], x_B[inds_B])
        if return_inds:
            return torch.cat((inds_A, inds_B), dim=-1)
        return torch.cat((x_A[inds_A], x_B[inds_B]), dim=-1)

    def visualize_warp(self, warp, certainty, im_A=None, im_B=None, im_A_path=None, im_B_path=None, device="cuda",
                       symmetric=True, save_path=None, unnormalize=False):
        import numpy as np
        H, W2, _ = warp.shape
        W = W2 // 2 if symmetric else W2
        if im_A is None:
            im_A, im_B = Image.open(im_A_path).convert("RGB"), Image.open(im_B_path).convert("RGB")
        if not isinstance(im_A, torch.Tensor):
            im_A, im_B = im_A.resize((W, H)), im_B.resize((W, H))
            x_B = (torch.tensor(np.array(im_B)) / 255).to(device).permute(2, 0, 1)
            x_A = (torch.tensor(np.array(im_A)) / 255).to(device).permute(2, 0, 1) if symmetric else None
        else:
            x_A, x_B = (im_A if symmetric else None), im_B
        im_A_transfer = F.grid_sample(x_B[None], warp[:, :W, 2:][None], mode="bilinear", align_corners=False)[0]
        if symmetric:
            im_B_transfer = F.grid_sample(x_A[None], warp[:, W:, :2][None], mode="bilinear", align_corners=False)[0]
            warp_im = torch.cat((im_A_transfer, im_B_transfer), dim=2)
            white_im = torch.ones((H, 2 * W), device=device)
        else:
            warp_im, white_im = im_A_transfer, torch.ones((H, W), device=device)
        vis_im = certainty * warp_im + (1 - certainty) * white_im
        if save_path is not None:
            arr = vis_im
            if unnormalize:
                mean = torch.tensor([0.485, 0.456, 0.406], device=arr.device)[:, None, None]
                std = torch.tensor([0.229, 0.224, 0.225], device=arr.device)[:, None, None]
                arr = arr * std + mean
            arr = (arr.clamp(0, 1) * 255).byte().permute(1, 2, 0).cpu().numpy()
            Image.fromarray(arr).save(save_path)
        return vis_im
