"""GPU: the CUDA path against the committed golden fixtures that were produced by running the reference."""
import os

import numpy as np
import pytest
import torch

from turkish_asr_model_b200 import _lib as L
from turkish_asr_model_b200.data.preprocessing import AudioPreprocessor
from turkish_asr_model_b200.model import TurkishASRModel
from turkish_asr_model_b200.utils.decoding import GreedyDecoder

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_mel_vs_reference_golden(cuda):
    g = np.load(os.path.join(GOLD, "mel_golden.npz"))
    pre = AudioPreprocessor(device="cuda")
    lengths = g["lengths"].tolist()
    waves = torch.zeros(len(lengths), max(lengths))
    for i, n in enumerate(lengths):
        waves[i, :n] = torch.from_numpy(g["wave%d" % i])
    feats, frames = pre.extract_features_batch(waves.to(cuda), torch.tensor(lengths))
    for i, n in enumerate(lengths):
        ref = g["feat%d" % i]
        assert int(frames[i]) == ref.shape[0]  # integer parity
        assert np.abs(feats[i, : ref.shape[0]].cpu().numpy() - ref).max() <= 1e-4
        assert torch.all(feats[i, ref.shape[0]:] == 0)
        single = pre.extract_features(torch.from_numpy(g["wave%d" % i]))  # reference call signature
        assert single.shape == ref.shape and np.abs(single.cpu().numpy() - ref).max() <= 1e-4


def test_model_vs_reference_golden(cuda):
    g = np.load(os.path.join(GOLD, "model_golden.npz"))
    torch.manual_seed(0)
    model = TurkishASRModel(80, 128, 2, 1, 32, dropout=0.0)
    for name, (s, a) in zip(g["checksum_names"], g["checksums"]):  # identical seeded init as the reference
        assert abs(float(model.state_dict()[str(name)].double().sum()) - s) <= 1e-9 * max(1.0, abs(a))
    model = model.to(cuda).train()
    x = torch.from_numpy(g["x"]).to(cuda)
    il = torch.from_numpy(g["input_lengths"])
    logits = model(x, il)
    ref = torch.from_numpy(g["logits_train"])
    rel = ((logits.float().cpu() - ref).abs().max() / ref.abs().max()).item()
    assert rel < 2e-2  # north_star: encoder logits 2e-2 relative in bf16
    loss, _, dlogits = L.ctc_loss_fwd_bwd(logits.detach(), torch.from_numpy(g["targets"]).to(cuda), (il // 4).to(cuda),
                                          torch.from_numpy(g["target_lengths"]).to(cuda))
    assert abs(loss.item() - float(g["loss"])) < 2e-2 * abs(float(g["loss"]))
    logits.backward(dlogits)
    grads = dict(model.named_parameters())
    rels = []
    for name, ref_norm in zip(g["grad_names"], g["grad_norms"]):
        p = grads[str(name)]
        if ref_norm < 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        elif not str(name).endswith("depthwise_conv.bias"):
            rels.append(abs(float(p.grad.norm()) - ref_norm) / max(ref_norm, 1e-8))
    assert max(rels) < 6e-2 and np.median(rels) < 1.5e-2
    # eval-mode logits (BatchNorm running statistics updated by the train-mode pass above) and greedy ids
    model.eval()
    with torch.no_grad():
        le = model(x, il)
    ref_e = torch.from_numpy(g["logits_eval"])
    assert ((le.float().cpu() - ref_e).abs().max() / ref_e.abs().max()).item() < 2e-2
    ids, _, _ = L.argmax_collapse(ref_e.to(cuda).contiguous())
    assert np.array_equal(ids.cpu().numpy(), g["greedy_ids"])  # bit-exact on identical logits
    dec = GreedyDecoder(None, blank_id=0).decode_ids_batch(ref_e.to(cuda).contiguous())
    for b in range(2):
        seq, last = [], None
        for c in g["greedy_ids"][b].tolist():
            if c != last and c != 0:
                seq.append(c)
            last = c
        assert dec[b] == seq
