"""Restatement of the reference's representation producers -- TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/code/model_2D.py:20-55`` (``FeatureExtractor``), ``train_arco_2d.py:231-236``
(``q_representation``, the two extractors) and the composition at ``train_arco_2d.py:313-333`` that turns the decoder's five
feature maps into ``rep`` / ``rep_teacher``.  Functional (weights passed in) so the same code serves CPU and GPU checks.
Pinned by ``tests/golden/producers_*.npz``, made by ``tests/golden/make_golden_producers.py`` from the reference's own
class / statement source.
"""
import torch
import torch.nn.functional as F

from .contra_oracle import contra_memobank_loss


def conv1x1(x, w):
    """``nn.Conv2d(cin, cout, kernel_size=1, bias=False)`` with weight ``w`` [cout, cin] or [cout, cin, 1, 1]."""
    return F.conv2d(x, w.reshape(w.shape[0], w.shape[1], 1, 1))


def _up(x, ref):
    # nn.Upsample(size=ref.shape[-2:], mode='bilinear', align_corners=True)   (model_2D.py:43,46,49,52)
    return F.interpolate(x, size=ref.shape[-2:], mode="bilinear", align_corners=True)


def feature_extractor_trunk(w, fea_list):
    """``FeatureExtractor.forward`` up to the input of ``fea4`` (model_2D.py:36-53); ``w = [fea0 .. fea4]`` weights."""
    f0, f1, f2, f3, f4 = fea_list[:5]
    x = conv1x1(f0, w[0]) + f0                      # :42
    x = torch.cat((_up(x, f1), f1), dim=1)          # :43-44
    x = conv1x1(x, w[1]) + x                        # :45
    x = torch.cat((_up(x, f2), f2), dim=1)          # :46-47
    x = conv1x1(x, w[2]) + x                        # :48
    x = torch.cat((_up(x, f3), f3), dim=1)          # :49-50
    x = conv1x1(x, w[3]) + x                        # :51
    return torch.cat((_up(x, f4), f4), dim=1)       # :52-53


def feature_extractor(w, fea_list):
    return conv1x1(feature_extractor_trunk(w, fea_list), w[4])     # :54


def q_representation(wq, x):
    """train_arco_2d.py:231-234: two bias-free 1x1 convolutions, no non-linearity in between."""
    return conv1x1(conv1x1(x, wq[0]), wq[1])


def representations(w_q_fe, w_q_rep, w_k_fe, maps_l, maps_u, maps_l_teacher, maps_u_teacher):
    """train_arco_2d.py:317-333: (rep_all, rep_teacher_all) from the decoder feature maps of the labelled / unlabelled
    batch through the student (q) and teacher (k) extractors."""
    rep_l = q_representation(w_q_rep, feature_extractor(w_q_fe, maps_l))        # :317,325
    rep_u = q_representation(w_q_rep, feature_extractor(w_q_fe, maps_u))        # :318,324
    rep_l_t = feature_extractor(w_k_fe, maps_l_teacher)                         # :321,329
    rep_u_t = feature_extractor(w_k_fe, maps_u_teacher)                         # :322,328
    return torch.cat((rep_l, rep_u)), torch.cat((rep_l_t, rep_u_t))             # :330,333


def contra_from_features(x_student, x_teacher, student_weights, teacher_weight, label_l, label_u, prob_l, prob_u, low_mask,
                         high_mask, memobank, queue_ptrlis, queue_size, **kw):
    """What ``arco_b200.producers.compute_contra_memobank_loss_from_features`` must equal: the reference composition
    ``loss(rep = chain(x_student), rep_teacher = conv(x_teacher))`` with both tensors materialised."""
    rep = x_student
    for w in student_weights:
        rep = conv1x1(rep, w.to(rep.dtype))
    rep_t = conv1x1(x_teacher, teacher_weight.to(x_teacher.dtype)).detach()
    # bf16 features: every convolution output is rounded to bf16 (autocast); the loss itself is evaluated in fp32 on those values
    rep, rep_t = rep.float(), rep_t.float()
    return contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, memobank, queue_ptrlis,
                                queue_size, rep_t, **kw)
