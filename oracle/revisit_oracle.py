"""Restatement of the reference's revisiting loss and random-pool queue -- TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/code/train_arco_2d.py``: ``get_revisiting_loss`` :126-136, ``_dequeue_and_enqueue`` :109-120
and the call site :400-402.  Pinned by ``tests/golden/revisit_*.npz`` (outputs of the reference's own function source,
executed by ``tests/golden/make_golden_step.py``).
"""
import torch
import torch.nn.functional as F


def revisiting_loss(random_pool, rep_u, rep_u_teacher, topk=5):
    """(:126-136) returns (loss, nn_index [bs, topk], dist_t, dist_q)."""
    s = F.normalize(rep_u.reshape(rep_u.shape[0], -1).float(), dim=-1)                 # :127-128
    t = F.normalize(rep_u_teacher.reshape(rep_u_teacher.shape[0], -1).float(), dim=-1)  # :129-130
    dist_t = 2 - 2 * torch.einsum("bc,kc->bk", s, random_pool)                         # :131  (from the student)
    dist_q = 2 - 2 * torch.einsum("bc,kc->bk", t, random_pool)                         # :132  (from the teacher)
    _, nn_index = dist_t.topk(topk, dim=1, largest=False)                              # :133
    nn_dist_q = torch.gather(dist_q, 1, nn_index)                                      # :134
    loss = (nn_dist_q.sum(dim=1) / topk).mean()                                        # :135
    return loss, nn_index, dist_t, dist_q


def pool_enqueue(rep_u_teacher, random_pool, random_pool_ptr):
    """(:400-402 + :109-120) in place on ``random_pool`` / ``random_pool_ptr``."""
    keys = F.normalize(rep_u_teacher.reshape(rep_u_teacher.shape[0], -1).float(), dim=-1)
    bs, K = keys.shape[0], random_pool.shape[0]
    ptr = int(random_pool_ptr)
    assert K % bs == 0
    random_pool[ptr: ptr + bs] = keys
    random_pool_ptr[0] = (ptr + bs) % K
