"""Offline factories for synthetic runs: tokenizer, processor and random-init models of the
architectures the reference attacks (utils/experiments_utils.py:74-138 loads the real checkpoints;
there is no network here, so weights are random and prompts are synthetic token strings).

Nothing in this module is on the timed path; it only manufactures the objects that
``bimodalattack_b200.run`` (and the reference's ``bimodalattack.run``) take as arguments.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

CLIP_MEAN = [0.48145466, 0.4578275, 0.40821073]     # reference experiments.py:394-397
CLIP_STD = [0.26862954, 0.26130258, 0.27577711]
GEMMA_MEAN = [0.5, 0.5, 0.5]                         # reference experiments.py:383
GEMMA_STD = [0.5, 0.5, 0.5]

SPECIALS = ["<pad>", "<s>", "</s>", "<unk>"]


def make_tokenizer(vocab_size: int, image_token: str = "<image>", extra_tokens: Sequence[str] = ()):
    """Word-level tokenizer whose decode->encode round trip is the identity, so ``filter_ids``
    keeps every candidate (reference bimodal_attack.py:166-186). ids: 0..3 specials, 4 = image
    token, then extra tokens, then ``x``, then w0, w1, ..."""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast

    words = list(SPECIALS) + [image_token] + list(extra_tokens) + ["x", "USER:", "ASSISTANT:"]
    n_fill = vocab_size - len(words)
    if n_fill < 0:
        raise ValueError("vocab_size too small")
    words += [f"w{i}" for i in range(n_fill)]
    vocab = {w: i for i, w in enumerate(words)}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<s>", eos_token="</s>",
                                   unk_token="<unk>", pad_token="<pad>")
    return fast


class Gemma3Processor:  # noqa: N801 - the reference dispatches on this exact class name (:314)
    """Minimal stand-in for HF's Gemma3Processor: the reference only reads the class name, the
    ``chat_template`` attribute and calls ``apply_chat_template``."""

    def __init__(self, tokenizer, generation_prompt: str = " <end_of_turn> <start_of_turn>"):
        self.tokenizer = tokenizer
        self.chat_template = "gemma3-synthetic"
        self.generation_prompt = generation_prompt

    def apply_chat_template(self, messages, add_generation_prompt=True, **_):
        content = messages[-1]["content"]
        if isinstance(content, str):
            text, has_image = content, False
        else:
            text = " ".join(c["text"] for c in content if c.get("type") == "text")
            has_image = any(c.get("type") == "image" for c in content)
        out = "<start_of_turn> " + text
        if has_image:
            out += " <start_of_image>"
        out += self.generation_prompt
        return out


def make_llava_processor(tokenizer, image_size: int, patch_size: int = 14):
    from transformers import CLIPImageProcessor, LlavaProcessor

    ip = CLIPImageProcessor(size={"shortest_edge": image_size},
                            crop_size={"height": image_size, "width": image_size})
    return LlavaProcessor(image_processor=ip, tokenizer=tokenizer, patch_size=patch_size,
                          vision_feature_select_strategy="default")


def _llava_config(vocab, d, layers, heads, kv_heads, d_ff, head_dim, v_d, v_layers, v_heads, v_ff,
                  image_size, patch, image_token_index=4, max_pos=4096):
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig

    vc = CLIPVisionConfig(hidden_size=v_d, intermediate_size=v_ff, num_hidden_layers=v_layers,
                          num_attention_heads=v_heads, image_size=image_size, patch_size=patch,
                          projection_dim=v_d)
    tc = LlamaConfig(vocab_size=vocab, hidden_size=d, intermediate_size=d_ff,
                     num_hidden_layers=layers, num_attention_heads=heads,
                     num_key_value_heads=kv_heads, head_dim=head_dim,
                     max_position_embeddings=max_pos, rms_norm_eps=1e-5, rope_theta=10000.0,
                     tie_word_embeddings=False)
    return LlavaConfig(vision_config=vc, text_config=tc, image_token_index=image_token_index,
                       vision_feature_layer=-2, vision_feature_select_strategy="default",
                       tie_word_embeddings=False)


def _finish(model, dtype, device, head_gain: float, seed: int):
    """HF's N(0, 0.02) init gives near-uniform logits (every candidate's bf16 loss is identical,
    SURVEY.md section 7 'Ties'); scaling the LM head by ``head_gain`` spreads the per-candidate
    losses so argmin parity means something. Seeded and identical on every rank."""
    with torch.no_grad():
        head = model.get_output_embeddings()
        emb = model.get_input_embeddings()
        if head is not None and head.weight is not emb.weight:
            head.weight.mul_(head_gain)
    model = model.to(dtype=dtype)
    if device is not None:
        model = model.to(device)
    return model.eval().requires_grad_(False)


def _tiled_init(model, seed: int):
    """Fast CPU initialisation for timing-only models: every matrix is tiled from one N(0, 0.02)
    block of 2^20 values, norm weights are 1, biases 0 (HF's per-parameter normal_ takes minutes
    for 7B parameters on a few cores)."""
    g = torch.Generator().manual_seed(seed)
    block = torch.randn(1 << 20, generator=g) * 0.02
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() < 2:
                p.fill_(1.0) if ("norm" in name and name.endswith("weight")) else p.zero_()
                continue
            flat = p.view(-1)
            reps = (flat.numel() + block.numel() - 1) // block.numel()
            flat.copy_(block.repeat(reps)[: flat.numel()].to(p.dtype))


def make_llava(size: str = "tiny", dtype=torch.float32, device=None, seed: int = 0,
               head_gain: float = 20.0, attn_implementation: Optional[str] = None,
               fast_init: bool = False):
    """size: 'tiny' (CPU config of BASELINE.json: 2 layers, d=256, 4-token image),
    'tiny128' (same but head_dim 128 so the CUDA scorer accepts it), '7b' (LLaVA-1.5-7B arch)."""
    from transformers import LlavaForConditionalGeneration

    if size == "tiny":
        cfg = _llava_config(1000, 256, 2, 4, 4, 512, 64, 64, 3, 4, 128, 28, 14)
    elif size == "tiny128":
        cfg = _llava_config(1024, 256, 2, 2, 2, 512, 128, 64, 3, 4, 128, 28, 14)
    elif size == "small128":
        cfg = _llava_config(4096, 512, 3, 4, 4, 1024, 128, 128, 3, 4, 256, 56, 14)
    elif size == "small128v":      # small128 with a CLIP tower of head_dim 64 (what the native tower covers), 4 layers
        cfg = _llava_config(4096, 512, 3, 4, 4, 1024, 128, 128, 4, 2, 256, 56, 14)
    elif size == "7b":
        cfg = _llava_config(32064, 4096, 32, 32, 32, 11008, 128, 1024, 24, 16, 4096, 336, 14)
    else:
        raise ValueError(size)
    if attn_implementation is not None:
        cfg._attn_implementation = attn_implementation
    torch.manual_seed(seed)
    if fast_init:
        from transformers.initialization import no_init_weights
        with no_init_weights():
            model = LlavaForConditionalGeneration(cfg)
        _tiled_init(model, seed)
    elif device is not None and torch.device(device).type == "cuda":
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with torch.device(device):
                model = LlavaForConditionalGeneration(cfg)
        finally:
            torch.set_default_dtype(prev)
    else:
        model = LlavaForConditionalGeneration(cfg)
    return _finish(model, dtype, device, head_gain, seed)


def swap_in_robust_clip(llava, image_size: int = 224, seed: int = 7):
    """LLaVA-RC: a second CLIP ViT tower of the same width swapped into the model (reference
    experiments_utils.py:103-136, RCLIP ViT-L at 224 px). On transformers >= 5 the tower lives at
    ``llava.model.vision_tower``."""
    from transformers import CLIPVisionConfig, CLIPVisionModel

    old = llava.model.vision_tower
    oc = old.config
    vc = CLIPVisionConfig(hidden_size=oc.hidden_size, intermediate_size=oc.intermediate_size,
                          num_hidden_layers=oc.num_hidden_layers,
                          num_attention_heads=oc.num_attention_heads, image_size=image_size,
                          patch_size=oc.patch_size, projection_dim=oc.projection_dim)
    p = next(old.parameters())
    torch.manual_seed(seed)
    tower = CLIPVisionModel(vc).to(device=p.device, dtype=p.dtype).eval().requires_grad_(False)
    llava.model.vision_tower = tower
    return llava


def make_gemma3(size: str = "tiny", dtype=torch.float32, device=None, seed: int = 0,
                head_gain: float = 1.0, embed_gain: float = 1.0, randomize_norms: bool = False,
                fast_init: bool = False):
    """Gemma-3 (SigLIP tower, tied LM head). Random init leaves the multimodal projector at zero
    (the image would have no effect), so it is re-initialised here. ``embed_gain`` scales the tied
    embedding / LM-head table (spreads the per-candidate losses like ``head_gain`` does for an
    untied head); ``randomize_norms`` gives the (1 + w) RMSNorm weights non-zero values.
    Sizes: 'tiny' (head_dim 128, CPU golden cases), 'tiny256' (head_dim 256, GQA 2, one local and
    one global layer: the 4b attention geometry in miniature), '4b' (Gemma-3-4b-it arch)."""
    from transformers import Gemma3Config, Gemma3ForConditionalGeneration, Gemma3TextConfig, SiglipVisionConfig

    if size == "tiny":
        tc = Gemma3TextConfig(vocab_size=1024, hidden_size=256, intermediate_size=512,
                              num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                              head_dim=128, max_position_embeddings=2048, sliding_window=1024,
                              query_pre_attn_scalar=128)
        vc = SiglipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                num_attention_heads=4, image_size=56, patch_size=14,
                                vision_use_head=False)
        cfg = Gemma3Config(text_config=tc, vision_config=vc, mm_tokens_per_image=4,
                           image_token_index=4, boi_token_index=5, eoi_token_index=6)
    elif size == "tiny256":
        tc = Gemma3TextConfig(vocab_size=1024, hidden_size=256, intermediate_size=512,
                              num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1,
                              head_dim=256, max_position_embeddings=2048, sliding_window=1024,
                              query_pre_attn_scalar=256,
                              layer_types=["sliding_attention", "full_attention"])
        vc = SiglipVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                                num_attention_heads=4, image_size=56, patch_size=14,
                                vision_use_head=False)
        cfg = Gemma3Config(text_config=tc, vision_config=vc, mm_tokens_per_image=4,
                           image_token_index=4, boi_token_index=5, eoi_token_index=6)
    elif size == "4b":
        tc = Gemma3TextConfig(vocab_size=262208, hidden_size=2560, intermediate_size=10240,
                              num_hidden_layers=34, num_attention_heads=8, num_key_value_heads=4,
                              head_dim=256, max_position_embeddings=131072, sliding_window=1024,
                              query_pre_attn_scalar=256, rope_theta=1000000.0,
                              rope_local_base_freq=10000.0)
        vc = SiglipVisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                                num_attention_heads=16, image_size=896, patch_size=14,
                                vision_use_head=False)
        cfg = Gemma3Config(text_config=tc, vision_config=vc, mm_tokens_per_image=256,
                           image_token_index=262144, boi_token_index=255999, eoi_token_index=256000)
    else:
        raise ValueError(size)
    torch.manual_seed(seed)
    if fast_init:
        from transformers.initialization import no_init_weights
        with no_init_weights():
            model = Gemma3ForConditionalGeneration(cfg)
        _tiled_init(model, seed)
        with torch.no_grad():
            for name, p in model.named_parameters():      # (1 + w) norms: w = 0 is the identity scale
                if "norm" in name and p.dim() < 2:
                    p.zero_()
    elif device is not None and torch.device(device).type == "cuda":
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with torch.device(device):
                model = Gemma3ForConditionalGeneration(cfg)
        finally:
            torch.set_default_dtype(prev)
    else:
        model = Gemma3ForConditionalGeneration(cfg)
    with torch.no_grad():
        proj = model.model.multi_modal_projector
        torch.manual_seed(seed + 1)
        proj.mm_input_projection_weight.copy_(
            torch.randn(proj.mm_input_projection_weight.shape) * 0.02)
        if embed_gain != 1.0:
            model.get_input_embeddings().weight.mul_(embed_gain)
        if randomize_norms:
            g = torch.Generator().manual_seed(seed + 2)
            for name, p in model.model.language_model.named_parameters():
                if "norm" in name:
                    p.copy_((torch.rand(p.shape, generator=g) - 0.5).to(p.device, p.dtype))
    return _finish(model, dtype, device, head_gain, seed)


def synthetic_words(n: int, start: int = 10) -> str:
    return " ".join(f"w{start + i}" for i in range(n))


def headline_case(vocab: int = 32000, image_px: int = 336, n_optim: int = 20, n_target: int = 16,
                  seed: int = 1234):
    """Synthetic inputs of BASELINE.json config 2: segment lengths
    before_img 4 | image (px/14)^2 | before_suffix 16 | optim 20 | after 6 | target 16, i.e. a
    596-row shared prefix and a 42-row tail at 336 px. Returns (tokenizer, processor, normalize,
    image [cpu fp32], messages, target, optim_str_init)."""
    import torchvision.transforms as T

    tok = make_tokenizer(vocab)
    proc = make_llava_processor(tok, image_px)
    template = ("w1 w2 w3 <image> {{ messages[0]['content'][0]['text'] }} "
                + synthetic_words(6, start=100))
    tok.chat_template = template
    proc.chat_template = template
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((1, 3, image_px, image_px), generator=g)
    messages = [{"role": "user", "content": synthetic_words(15, start=200)}]
    target = synthetic_words(n_target, start=400)
    init = " ".join(["x"] * n_optim)
    return tok, proc, T.Normalize(CLIP_MEAN, CLIP_STD), image, messages, target, init


def gemma_case(vocab: int = 262208, image_px: int = 896, n_optim: int = 20, n_target: int = 16,
               seed: int = 1234):
    """Synthetic inputs of BASELINE.json config 4: Gemma-3 text-only order
    before 20 | optim 20 | after 6 | target 16. Returns (tokenizer, processor, normalize,
    image [cpu fp32], messages, target, optim_str_init)."""
    import torchvision.transforms as T

    tok = make_tokenizer(vocab, image_token="<image_soft_token>",
                         extra_tokens=["<start_of_image>", "<end_of_image>", "<start_of_turn>", "<end_of_turn>"])
    proc = Gemma3Processor(tok, generation_prompt=" <end_of_turn> <start_of_turn> " + synthetic_words(4, start=100))
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((1, 3, image_px, image_px), generator=g)
    messages = [{"role": "user", "content": synthetic_words(18, start=200)}]     # + <s> <start_of_turn> = 20
    target = synthetic_words(n_target, start=400)
    init = " ".join(["x"] * n_optim)
    return tok, proc, T.Normalize(GEMMA_MEAN, GEMMA_STD), image, messages, target, init
