"""GPU parity: the CLIP text executors (tcgen05 GEMMs with fused q|k|v, causal tcgen05 attention read in place from
the fused projection, LayerNorm kernels, quick-GELU folded into the GEMM epilogues) vs the HF fp32 modules on the same
random-init weights, and ``encode_prompt`` end to end on prompt strings vs the oracle restatement.

Tolerances: relative L2 error of 16-bit arithmetic over 12 - 32 pre-LN blocks; an fp16 / bf16 *emulation* of the same schedule
(tests/cpu_ops_emulation.py, rounding every op output to the storage type) gives 1.1e-3 / 9e-3 at these geometries."""
import pytest
import torch

from oracle import comat_ref as R
from oracle import fixtures as FX

pytestmark = pytest.mark.gpu

PROMPTS = ["a photo of a cat", "two red cubes on a blue sphere next to a green cone", "",
           "a b c d e f g h i j k l m n o p q r s t u v w x y z", " ".join("w%d" % i for i in range(100))]


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("which,layers,dtype,tol", [("clip_l", None, torch.float16, 5e-3), ("clip_l", None, torch.bfloat16, 3e-2),
                                                    ("bigg", 8, torch.float16, 5e-3), ("bigg", 32, torch.float16, 5e-3)])
def test_clip_text_executor_vs_hf(which, layers, dtype, tol):
    from comat_b200 import _lib
    from comat_b200.text_encoder import EngineCLIPText
    model = R.make_clip_text(which, tiny=False, seed=3, device="cuda", layers=layers)
    t = FX.ClipTokenizerStub()(PROMPTS)
    ids = t.input_ids.cuda()
    with torch.no_grad():
        ref = model(ids, output_hidden_states=True)
    enc = EngineCLIPText(model, dtype)
    l0 = _lib.LAUNCH_COUNT
    out = enc(ids, output_hidden_states=True)
    n_layers = model.config.num_hidden_layers
    assert _lib.LAUNCH_COUNT - l0 >= 7 * n_layers        # 2 LN + 4 GEMM + 1 attention per block, all ours
    assert out.last_hidden_state.dtype == torch.float32 and len(out.hidden_states) == n_layers + 1
    print(f"[measured] {which} L={n_layers} {dtype}: last {rel(out.last_hidden_state, ref.last_hidden_state):.2e} "
          f"h[-2] {rel(out.hidden_states[-2], ref.hidden_states[-2]):.2e}")
    assert rel(out.last_hidden_state, ref.last_hidden_state) < tol
    assert rel(out.hidden_states[-2], ref.hidden_states[-2]) < tol
    if which == "bigg":
        assert rel(out[0], ref.text_embeds) < tol
    else:
        assert rel(out.pooler_output, ref.pooler_output) < tol
    # padded keys masked on top of the causal mask
    mask = t.attention_mask.cuda()
    with torch.no_grad():
        ref_m = model(ids, attention_mask=mask).last_hidden_state
    out_m = enc(ids, attention_mask=mask).last_hidden_state
    for i in range(len(PROMPTS)):
        L = int(mask[i].sum())
        assert rel(out_m[i, :L], ref_m[i, :L]) < tol


def test_encode_prompt_sd15_on_strings_vs_oracle():
    from comat_b200.pipelines import TrainableSDPipeline
    from comat_b200.synthetic import SyntheticClipTokenizer, build_clip_text
    from comat_b200.text_encoder import EngineCLIPText
    model = build_clip_text("cuda", torch.float32, seed=7)
    pipe = TrainableSDPipeline.__new__(TrainableSDPipeline)
    TrainableSDPipeline.__init__(pipe, vae=None, unet=None, text_encoder=EngineCLIPText(model, torch.float16), tokenizer=SyntheticClipTokenizer())
    dev = torch.device("cuda")
    pe, npe = pipe.encode_prompt(PROMPTS[:2], dev, 2, True)

    class _Tok(FX.ClipTokenizerStub):                                       # oracle tokenizer, ids moved to the model's device
        def __call__(self, *a, **k):
            t = super().__call__(*a, **k)
            t.input_ids = t.input_ids.cuda()
            return t
    pe_ref, npe_ref, _ = R.encode_prompt_sd(model, _Tok(), PROMPTS[:2], 2, True)
    assert pe.shape == pe_ref.shape == (4, 77, 768) and pe.is_cuda
    print(f"[measured] sd15 encode_prompt: {rel(pe, pe_ref):.2e} {rel(npe, npe_ref):.2e}")
    assert rel(pe, pe_ref) < 5e-3 and rel(npe, npe_ref) < 5e-3
    null = pipe.encode_prompt("", dev, 4, False)[0]                          # training_script.py:519
    assert null.shape == (4, 77, 768) and rel(null[:1], npe_ref[:1]) < 5e-3


def test_encode_prompt_sdxl_on_strings_vs_oracle():
    from comat_b200.pipelines import TrainableSDXLPipeline
    from comat_b200.synthetic import SyntheticClipTokenizer
    from comat_b200.text_encoder import EngineCLIPText
    e1 = R.make_clip_text("clip_l", tiny=False, seed=11, device="cuda")
    e2 = R.make_clip_text("bigg", tiny=False, seed=12, device="cuda", layers=6)
    pipe = TrainableSDXLPipeline.__new__(TrainableSDXLPipeline)
    TrainableSDXLPipeline.__init__(pipe, vae=None, unet=None, text_encoder=EngineCLIPText(e1, torch.float16), tokenizer=SyntheticClipTokenizer(),
                                   text_encoder_2=EngineCLIPText(e2, torch.float16), tokenizer_2=SyntheticClipTokenizer(pad_token_id=0),
                                   force_zeros_for_empty_prompt=False)

    class _Tok(FX.ClipTokenizerStub):
        def __call__(self, *a, **k):
            t = super().__call__(*a, **k)
            t.input_ids = t.input_ids.cuda()
            return t
    ref = R.encode_prompt_sdxl(e1, e2, _Tok(), _Tok(pad_token_id=0), PROMPTS[:3], 1, True, force_zeros_for_empty_prompt=False)
    got = pipe.encode_prompt(PROMPTS[:3], device=torch.device("cuda"), num_images_per_prompt=1, do_classifier_free_guidance=True)
    assert got[0].shape == (3, 77, 2048) and got[2].shape == (3, 1280)
    print("[measured] sdxl encode_prompt:", ["%.2e" % rel(a, b) for a, b in zip(got, ref)])
    for a, b in zip(got, ref):
        assert rel(a, b) < 5e-3


def test_graphed_encoder_replays_follow_inputs_and_match_eager():
    """the captured forward (static id / mask buffers) returns what the eager launches return, call after call, per signature."""
    from comat_b200.text_encoder import EngineCLIPText
    model = R.make_clip_text("bigg", tiny=False, seed=5, device="cuda", layers=3)
    enc = EngineCLIPText(model, torch.float16)
    tok = FX.ClipTokenizerStub()
    a, b = tok(PROMPTS[:3]), tok(PROMPTS[2:5])
    enc.use_graphs = False
    want = [enc(t.input_ids.cuda(), output_hidden_states=True) for t in (a, b)]
    want_m = enc(a.input_ids.cuda(), attention_mask=a.attention_mask.cuda())
    enc.use_graphs = True
    for t, w in ((a, want[0]), (b, want[1]), (b, want[1]), (a, want[0])):
        got = enc(t.input_ids.cuda(), output_hidden_states=True)
        assert torch.equal(got.text_embeds, w.text_embeds) and torch.equal(got.last_hidden_state, w.last_hidden_state)
        assert all(torch.equal(x, y) for x, y in zip(got.hidden_states, w.hidden_states))
    got_m = enc(a.input_ids.cuda(), attention_mask=a.attention_mask.cuda())
    assert torch.equal(got_m.last_hidden_state, want_m.last_hidden_state)
    assert len(enc._graphs) == 2                                      # (3,77) with hidden states; (3,77) with a mask
    # results handed out earlier are copies, not views of the graph's static buffers
    first = enc(a.input_ids.cuda(), output_hidden_states=True).last_hidden_state
    keep = first.clone()
    enc(b.input_ids.cuda(), output_hidden_states=True)
    assert torch.equal(first, keep)
