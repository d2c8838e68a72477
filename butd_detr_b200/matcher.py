"""Hungarian matcher on the device: a drop-in for the reference's
`models.losses.HungarianMatcher` (`/root/reference/models/losses.py:232-331`) — same constructor, same
`forward(outputs, targets)` contract — whose cost matrix and assignment run on the sm_100a kernels
(`csrc/matcher.cu`) instead of torch launches + a device-to-host copy + scipy on the CPU:

    set_criterion = SetCriterion(matcher=butd_detr_b200.matcher.HungarianMatcher(1, 0, 2, True), ...)

The returned index tensors live on the device (the reference returns CPU tensors); every consumer in
`SetCriterion` (losses.py:353-505) indexes device tensors with them, so nothing waits for the host:
the matcher's 7 synchronisations per training step are gone.  There is no CPU path.
"""
import torch
from torch import nn

from . import _lib


class HungarianMatcher(nn.Module):
    """Assignment between targets and predictions (1-to-1, minimum total cost)."""

    def __init__(self, cost_class=1, cost_bbox=5, cost_giou=2, soft_token=False):
        super().__init__()
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0
        self.cost_class, self.cost_bbox, self.cost_giou, self.soft_token = cost_class, cost_bbox, cost_giou, soft_token
        self.last_cost = None    # (sum T_b, Q) target-major costs of the last call (device)
        self.last_status = None  # device int: 1 when a scene had no finite assignment (scipy would raise)

    @torch.no_grad()
    def forward(self, outputs, targets):
        """outputs: {"pred_logits" (B,Q,C), "pred_boxes" (B,Q,6) cxcyczwhd}; targets: per scene {"labels" (T),
        "boxes" (T,6), "positive_map" (T,256)}.  Returns [(index_i, index_j)] per scene: matched predictions
        (ascending) and their targets, int64, on the device."""
        logits = outputs["pred_logits"]
        _lib.check_cuda(logits)
        dev = logits.device
        logits = logits.detach().float().contiguous()
        boxes = outputs["pred_boxes"].detach().float().contiguous()
        B, Q, C = logits.shape
        sizes = [int(t["boxes"].shape[0]) for t in targets]  # host-known shapes: no device synchronisation
        assert len(sizes) == B
        if max(sizes, default=0) > Q:
            raise NotImplementedError("more targets than queries in a scene")
        total = sum(sizes)
        empty = torch.empty(0, dtype=torch.int64, device=dev)
        if total == 0:
            return [(empty, empty) for _ in range(B)]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        tgt_off = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
        tgt_boxes = torch.cat([t["boxes"] for t in targets]).float().contiguous()
        pm = labels = None
        if self.soft_token:
            pm = torch.cat([t["positive_map"] for t in targets]).float().contiguous()
            if pm.shape[-1] < C:
                raise ValueError("positive_map has fewer columns than there are classes")
        else:
            labels = torch.cat([t["labels"] for t in targets]).to(torch.int64).contiguous()
        cost = torch.empty(total, Q, dtype=torch.float32, device=dev)
        mq = torch.empty(total, dtype=torch.int64, device=dev)
        mt = torch.empty(total, dtype=torch.int64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("bd_matcher_cost", logits.data_ptr(), boxes.data_ptr(), tgt_boxes.data_ptr(), _lib.ptr(pm),
                      0 if pm is None else pm.stride(0), _lib.ptr(labels), tgt_off.data_ptr(), B, Q, C,
                      float(self.cost_class), float(self.cost_bbox), float(self.cost_giou), cost.data_ptr())
            _lib.call("bd_hungarian", cost.data_ptr(), tgt_off.data_ptr(), B, Q, max(sizes), mq.data_ptr(), mt.data_ptr(),
                      status.data_ptr())
        self.last_cost, self.last_status = cost, status
        return [(mq[offs[b]:offs[b + 1]], mt[offs[b]:offs[b + 1]]) for b in range(B)]
