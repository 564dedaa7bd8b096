"""TEST INFRASTRUCTURE -- CPU restatement of the trainers' mask / threshold preparation.

Follows ``/root/reference/code/train_arco_2d.py:345-393`` (3-D twin ``train_arco_3d.py:315-353``) line by line with the
same torch / numpy calls; pinned by ``tests/golden/prepare_*.npz``, which ``tests/golden/make_golden.py`` produces by
EXECUTING the reference's own lines (read from the reference tree at generation time, never copied here).
Only tests/, ``__graft_entry__.smoke()`` and bench.py's baseline legs may import this package.
"""
import numpy as np
import torch


def label_onehot(inputs: torch.Tensor, num_segments: int) -> torch.Tensor:
    """train_arco_2d.py:492-498 / train_arco_3d.py:463-469: relu(label) (ignore -1 -> class 0), float one-hot."""
    idx = torch.relu(inputs).detach().cpu().type(torch.int64)
    out = torch.zeros((inputs.shape[0], num_segments) + tuple(inputs.shape[1:]))
    return out.scatter_(1, idx.unsqueeze(1), 1.0)


def prepare(pred_u, pred_l_teacher, pred_u_teacher, train_l_label, train_u_aug_label, alpha_t, num_classes):
    with torch.no_grad():
        label_l = label_onehot(train_l_label, num_classes)                                  # :349 (interpolate = identity)
        label_u = label_onehot(train_u_aug_label, num_classes)                              # :350
        prob_u = torch.softmax(pred_u, dim=1)                                               # :353
        prob_l_teacher = torch.softmax(pred_l_teacher, dim=1)                               # :355
        prob_u_teacher = torch.softmax(pred_u_teacher, dim=1)                               # :356
        entropy = -torch.sum(prob_u * torch.log(prob_u + 1e-10), dim=1)                     # :357-358
        low, high, low_thresh, high_thresh = masks_from_entropy(entropy, train_l_label, train_u_aug_label, alpha_t)
    return dict(label_l=label_l, label_u=label_u, prob_l_teacher=prob_l_teacher, prob_u_teacher=prob_u_teacher,
                low_mask_all=low, high_mask_all=high, entropy=entropy,
                thresholds=np.asarray([low_thresh, high_thresh], np.float32))


def masks_from_entropy(entropy, train_l_label, train_u_aug_label, alpha_t):
    valid = train_u_aug_label >= 0
    low_thresh = np.percentile(entropy[valid].cpu().numpy().flatten(), alpha_t)             # :359-361
    low_entropy_mask = entropy.le(low_thresh).float() * valid.bool()                        # :362-364
    high_thresh = np.percentile(entropy[valid].cpu().numpy().flatten(), 100 - alpha_t)      # :365-368
    high_entropy_mask = entropy.ge(high_thresh).float() * valid.bool()                      # :369-371
    lab = (train_l_label.unsqueeze(1) >= 0).float()
    low = torch.cat((lab, low_entropy_mask.unsqueeze(1)))                                   # :373-378
    high = torch.cat((lab, high_entropy_mask.unsqueeze(1)))                                 # :384-389
    return low, high, low_thresh, high_thresh
