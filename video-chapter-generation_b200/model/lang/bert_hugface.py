"""Mirror of the reference's model/lang/bert_hugface.py: the BERT-base text stream.

``BertHugface(pretrain_stage=False)`` exposes ``base_model`` (parameter container keyed like transformers'
``BertModel.state_dict()``), ``embed_size`` (768), ``vocab_size`` and ``build_chapter_head()`` exactly as
video_chapter_generation/model/lang/bert_hugface.py:13-36.  The reference's constructor downloads
'bert-base-uncased' (:20); offline, parameters start from BERT's initialiser (normal(0, 0.02), unit LayerNorm) and
are expected to come from ``load_state_dict``.  All arithmetic happens in libvcg_b200.so via TwoStream.forward.
"""
import torch
import torch.nn as nn


class _Holder(nn.Module):
    def forward(self, *a, **kw):
        raise RuntimeError("this module only holds parameters; the text stream runs inside libvcg_b200.so "
                           "(call TwoStream.forward)")


class _Embeddings(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.word_embeddings = nn.Embedding(cfg.vocab_size, cfg.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(cfg.max_position_embeddings, cfg.hidden_size)
        self.token_type_embeddings = nn.Embedding(cfg.type_vocab_size, cfg.hidden_size)
        self.LayerNorm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class _SelfAttention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.query = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.key = nn.Linear(cfg.hidden_size, cfg.hidden_size)
        self.value = nn.Linear(cfg.hidden_size, cfg.hidden_size)


class _DenseNorm(_Holder):
    def __init__(self, n_in, n_out, eps):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)
        self.LayerNorm = nn.LayerNorm(n_out, eps=eps)


class _Dense(_Holder):
    def __init__(self, n_in, n_out):
        super().__init__()
        self.dense = nn.Linear(n_in, n_out)


class _Attention(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.self = _SelfAttention(cfg)
        self.output = _DenseNorm(cfg.hidden_size, cfg.hidden_size, cfg.layer_norm_eps)


class _Layer(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg)
        self.intermediate = _Dense(cfg.hidden_size, cfg.intermediate_size)
        self.output = _DenseNorm(cfg.intermediate_size, cfg.hidden_size, cfg.layer_norm_eps)


class _Encoder(_Holder):
    def __init__(self, cfg):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(cfg) for _ in range(cfg.num_hidden_layers)])


class BertBaseConfig:
    """bert-base-uncased geometry (transformers configuration_bert.py defaults)."""
    vocab_size = 30522
    hidden_size = 768
    num_hidden_layers = 12
    num_attention_heads = 12
    intermediate_size = 3072
    max_position_embeddings = 512
    type_vocab_size = 2
    layer_norm_eps = 1e-12
    initializer_range = 0.02


class BertParams(_Holder):
    """embeddings / encoder.layer.N / pooler, keyed like transformers.BertModel."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config or BertBaseConfig()
        self.embeddings = _Embeddings(self.config)
        self.encoder = _Encoder(self.config)
        self.pooler = _Dense(self.config.hidden_size, self.config.hidden_size)
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0.0, std)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, 0.0, std)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)


class BertHugface(nn.Module):
    def __init__(self, pretrain_stage=True):
        super().__init__()
        self.pretrain_stage = pretrain_stage
        self.base_model = BertParams()
        self.vocab_size = self.base_model.config.vocab_size
        self.embed_size = self.base_model.config.hidden_size
        self.head = nn.Linear(self.embed_size, self.vocab_size, bias=False)
        self.head.weight.data.normal_(mean=0.0, std=0.02)
        print("backbone's parameters: ", sum(p.numel() for p in self.base_model.parameters()))

    def build_chapter_head(self):
        self.head = nn.Linear(self.embed_size, 2)   # clip positive / negative (text-only mode)

    def fix_backbone(self):
        for pn, p in self.named_parameters():
            if "pooler" in pn or "head" in pn:
                continue
            p.requires_grad = False

    def forward(self, text_ids, attention_mask, get_attention=False):
        """(logits [B,2], prob [B,2]) = softmax(head(pooler_output)) — the clip classifier of --data_mode text
        (reference :98-132 with pretrain_stage=False).  BERT, pooler and head run inside libvcg_b200.so
        (VCG_MODALITY_TEXT engine); the MLM pre-training branch (pretrain_stage=True) is training code, out of scope."""
        import os
        if self.pretrain_stage:
            raise NotImplementedError("masked-LM pre-training (pretrain_stage=True) is out of scope (SURVEY.md section 2)")
        if not text_ids.is_cuda:
            raise RuntimeError("text-only scoring needs CUDA inputs: the B200 implementation has no CPU fallback")
        if self.training:
            raise RuntimeError("inference-only: call .eval() first")
        from vcg_b200.engine import Engine
        precision = getattr(self, "precision", os.environ.get("VCG_PRECISION", "bf16"))
        max_tokens = max(128, text_ids.shape[1])
        v = 0
        for t in list(self.parameters()) + list(self.buffers()):
            v += t._version + (t.data_ptr() % 1000003)
        key = (str(text_ids.device), precision, max_tokens, v)
        if getattr(self, "_engine", None) is None or self._engine_key != key:
            if getattr(self, "_engine", None) is not None:
                self._engine.close()
            eng = Engine(1, "mlp", precision, False, max_tokens, 256, 128, 8, device=text_ids.device, modality="text")
            eng.load_state_dict(self.state_dict())
            object.__setattr__(self, "_engine", eng)
            object.__setattr__(self, "_engine_key", key)
        with torch.no_grad():
            return self._engine.forward_text(text_ids, attention_mask)
