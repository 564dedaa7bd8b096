"""GPU: the config-5 dense tensor-core similarity (arco_similarity_dense) and its backward (arco_similarity_dense_backward)
against a plain PyTorch fp32 reference of the same op and its autograd gradient -- cos(anchor_q, bank[idx[q, n]]) as loss_helper_3d.py:466-486 forms it -- at 1e-5, including a wrapped ring,
a ragged last ring tile and a feature size that is not a multiple of the 64-element K block."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,Q,N,caps,fill", [
    (64, 128, 16, [300, 257], [300, 200]),          # ragged ring tiles, partly filled class
    (496, 256, 32, [1000, 512], [1000, 512]),       # the trainer's D: 7.75 K blocks (OOB columns zero-filled)
    (72, 128, 8, [130, 700], [130, 650]),
])
def test_dense_similarity_matches_torch(D, Q, N, caps, fill):
    from arco_b200.bank import DeviceMemoryBank
    from arco_b200.similarity import dense_similarity
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    rows = [torch.randn(fill[c], D, generator=g).to(torch.bfloat16).to(torch.float32) for c in range(2)]
    memobank = [[r.clone()] for r in rows]
    ptr = [torch.zeros(1, dtype=torch.long) for _ in range(2)]
    bank = DeviceMemoryBank(memobank, ptr, caps, D, dev, prefer_bf16=True)
    assert bank.row_dtype == torch.bfloat16
    # rotate class 0 so logical row r lives at (head + r) % cap
    bank.head[0] = 37
    shifted = torch.roll(bank.rows[: caps[0]].clone(), 37, dims=0)
    bank.rows[: caps[0]] = shifted
    anchors = (torch.randn(2, Q, D, generator=g) * 3).to(dev).requires_grad_(True)
    slot_classes = [1, 0]
    idx = torch.stack([torch.randint(0, fill[c], (Q, N), generator=g) for c in slot_classes]).to(torch.int32).to(dev)
    idx[:, :, 1] = idx[:, :, 0]                                                  # duplicates of a ring row accumulate in backward
    out = dense_similarity(anchors, bank, slot_classes, idx)
    g_out = torch.randn(out.shape, generator=g).to(dev)
    out.backward(g_out)
    a_ref = anchors.detach().clone().requires_grad_(True)
    for j, c in enumerate(slot_classes):
        keys = rows[c].to(dev)[idx[j].long()]                                    # [Q, N, D]
        ref = torch.nn.functional.cosine_similarity(a_ref[j][:, None, :], keys, dim=2)
        assert float((out[j].detach() - ref.detach()).abs().max()) <= 1e-5, (j, float((out[j].detach() - ref.detach()).abs().max()))
        (ref * g_out[j]).sum().backward()
    # the dense backward (second tcgen05 GEMM: scattered logit gradients x transposed ring) against autograd of the gather form
    err = float((anchors.grad - a_ref.grad).norm() / a_ref.grad.norm())
    assert err <= 1e-5, err
