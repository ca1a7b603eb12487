"""Drop-in for /root/reference/src/keypoints/grouping.py: ``MPPEHeatmapParser``.

Same constructor, same method names, same argument meaning, same return types (NumPy arrays where
the reference returns NumPy, tensors where it returns tensors) -- but every method runs the
sm_100a kernels of libhpdecode.so through ``torch.ops.hpd.*``.  Tensors that arrive on the CPU
are moved to the parser's CUDA device first (the reference accepts either); if no CUDA device or
no library is present the constructor raises: there is no CPU fallback.
"""
import numpy as np
import torch

from . import ops
from .decoder import _finish


class MPPEHeatmapParser(object):
    joints_order = list(ops.JOINTS_ORDER_17)   # grouping.py:63-65

    def __init__(self, num_kpts: int, max_num_people: int = 30, det_thr: float = 0.1, tag_thr: float = 1.0,
                 device="cuda:0"):
        if not torch.cuda.is_available():
            raise ops._lib.HpdError("MPPEHeatmapParser (hpdecode) needs a CUDA device; there is no CPU fallback")
        ops._lib.lib()
        self.max_num_people = max_num_people
        self.num_kpts = num_kpts
        self.det_thr = det_thr
        self.tag_thr = tag_thr
        self.device = torch.device(device)

    # -- helpers ------------------------------------------------------------------------------
    def _dev(self, x, dtype=torch.float32) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if not x.is_cuda:
            x = x.to(self.device, non_blocking=True)
        return x.to(dtype).contiguous()

    def _maps(self, kpts_hms, tags_hms):
        hm = self._dev(kpts_hms)
        tg = self._dev(tags_hms)
        if tg.dim() == 3:
            tg = tg.unsqueeze(-1)
        return hm.unsqueeze(0), tg.unsqueeze(0)

    # -- reference API ---------------------------------------------------------------------------
    def nms(self, kpts_heatmaps: torch.Tensor) -> torch.Tensor:
        """grouping.py:80-83 on [N,K,H,W]."""
        return torch.ops.hpd.nms(self._dev(kpts_heatmaps))

    def top_k(self, kpts_hms: torch.Tensor, tags_hms: torch.Tensor):
        """grouping.py:147-170: (tags_k f32[K,M,E], coords_k i32[K,M,2], scores_k f32[K,M]) as NumPy."""
        hm, tg = self._maps(kpts_hms, tags_hms)
        tags_k, coords_k, scores_k, _ = torch.ops.hpd.topk(hm, tg, self.max_num_people)
        return tags_k[0].cpu().numpy(), coords_k[0].cpu().numpy().astype(np.int32), scores_k[0].cpu().numpy()

    def match_by_tag(self, tags_k: np.ndarray, coords_k: np.ndarray, scores_k: np.ndarray) -> np.ndarray:
        """grouping.py:85-145: f32[P,K,3+E] (P may be 0)."""
        t = self._dev(tags_k).unsqueeze(0)
        c = self._dev(coords_k, torch.int32).unsqueeze(0)
        s = self._dev(scores_k).unsqueeze(0)
        # H*W only gates the top-k regime check; match_by_tag itself is size independent
        poses, n_person, flags = torch.ops.hpd.group(t, c, s, float(self.det_thr), float(self.tag_thr), 1 << 12, 1 << 12)
        P = int(n_person[0].item())
        if int(flags[0].item()) & 1:
            P = 0   # the reference's match_by_tag returns an empty array; parse() builds the fallback
        return poses[0, :P].cpu().numpy()

    def _adjust_refine(self, kpts_hms, tags_hms, grouped: np.ndarray, adjust: bool, refine: bool):
        hm, tg = self._maps(kpts_hms, tags_hms)
        _, K, H, W = hm.shape
        E = tg.shape[4]
        M = self.max_num_people
        bufs = ops.DecodeBuffers(1, K, H, W, E, M, hm.device, hm, tg)
        params = ops.make_params(1, K, H, W, E, M, self.det_thr, self.tag_thr, adjust, refine)
        ops.run_stage("nms", bufs, params)
        ops.run_stage("topk", bufs, params)
        P = grouped.shape[0]
        det = grouped[..., 2] != 0
        if det.any() and not ((grouped[..., 0][det] >= 0).all() and (grouped[..., 0][det] < W).all() and
                              (grouped[..., 1][det] >= 0).all() and (grouped[..., 1][det] < H).all()):
            raise IndexError("joint coordinates outside the %dx%d heatmap" % (H, W))   # the reference's indexing raises too
        bufs.poses.zero_()
        bufs.poses[0, :P] = torch.from_numpy(np.ascontiguousarray(grouped, dtype=np.float32)).to(hm.device)
        bufs.n_person.fill_(P)
        bufs.flags.zero_()
        ops.run_stage("adjust_refine", bufs, params)
        return bufs.poses[0, :P].cpu().numpy(), bufs.person_scores[0, :P].cpu().numpy()

    def adjust(self, grouped_joints: np.ndarray, kpts_hms: np.ndarray) -> np.ndarray:
        """grouping.py:172-191 (in place on grouped_joints, like the reference)."""
        K, H, W = kpts_hms.shape[-3:]
        E = grouped_joints.shape[-1] - 3
        dummy_tags = torch.zeros((K, H, W, E), device=self.device)
        out, _ = self._adjust_refine(kpts_hms, dummy_tags, grouped_joints, True, False)
        grouped_joints[...] = out
        return grouped_joints

    def refine(self, kpts_hms: np.ndarray, tags_hms: np.ndarray, person_joints: np.ndarray) -> np.ndarray:
        """grouping.py:193-250 for one person [K,3+E] (in place)."""
        out, _ = self._adjust_refine(kpts_hms, tags_hms, person_joints[None], False, True)
        person_joints[...] = out[0]
        return person_joints

    def parse(self, kpts_hms: torch.Tensor, tags_hms: torch.Tensor, adjust: bool = True, refine: bool = True):
        """grouping.py:252-283: (grouped_joints [P,K,3+E], person_scores [P])."""
        hm, tg = self._maps(kpts_hms, tags_hms)
        poses, scores, n_person, flags, *_ = torch.ops.hpd.parse(hm, tg, self.max_num_people, float(self.det_thr),
                                                                 float(self.tag_thr), adjust, refine)
        P = int(n_person[0].item())
        return _finish(poses[0, :P].cpu().numpy(), scores[0, :P].cpu().numpy(), int(flags[0].item()) & 1)
