"""GPT-2 inference engine on CPU: cached prefill + incremental decode must reproduce the
full-recompute forward (the reference's per-token strategy, examples/gpt2_inference.cpp:71-121)."""
import torch

from tnn_b200.inference import GPT2InferenceEngine
from tnn_b200.models import ExampleModels
from tnn_b200.nn import graph_from_sequential


def _tiny(layers_seed=3):
    from tnn_b200.models.zoo import _gpt2
    seq = _gpt2(64, 4, 2, False, "tiny_gpt", None, seq_len=32, vocab=97, dropout=0.1)
    return graph_from_sequential(seq, "cpu", layers_seed, "tiny_gpt")


def test_cached_decode_equals_full_recompute():
    g = _tiny()
    eng = GPT2InferenceEngine(g, max_batch=3)
    gen = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 97, (3, 9), generator=gen, dtype=torch.int32)
    nxt, logits = eng.prefill(ids, return_logits=True)
    full = eng.full_forward_logits(ids)
    torch.testing.assert_close(logits, full[:, -1, :], atol=1e-4, rtol=1e-4)
    assert torch.equal(nxt.long(), full[:, -1, :].argmax(-1))
    seq = ids
    tok = nxt
    for _ in range(5):
        seq = torch.cat([seq, tok.reshape(-1, 1)], 1)
        tok, lg = eng.decode(tok, return_logits=True)
        ref = eng.full_forward_logits(seq)[:, -1, :]
        torch.testing.assert_close(lg, ref, atol=2e-4, rtol=1e-3)
        assert torch.equal(tok.long(), ref.argmax(-1))
    assert eng.cur_len == 9 + 5


def test_generate_stops_at_eos_and_respects_capacity():
    g = _tiny(5)
    eng = GPT2InferenceEngine(g, max_batch=2, max_len=16)
    ids = torch.randint(0, 97, (2, 6), dtype=torch.int32)
    out = eng.generate(ids, max_new_tokens=8, eos_token=None)
    assert out.shape == (2, 8)
    first = int(out[0, 0])
    out2 = eng.generate(ids[:1].repeat(2, 1), max_new_tokens=8, eos_token=int(eng.prefill(ids[:1].repeat(2, 1))[0]))
    assert out2.shape[1] == 1                                    # first generated token is EOS for every row
    out3 = eng.generate(ids, max_new_tokens=100, eos_token=None)
    assert out3.shape[1] <= 16 - 6 + 1
    assert first == int(eng.generate(ids, 1, None)[0, 0])
