"""Drop-in for the reference `model.model` on the IISAN(Cached) path: same class names, constructor
arguments, forward signature, parameter names/shapes and registration order
(Code_Cached/model/model.py:15-105 ModelMM, :257-349 IISANAdaptedMModel; the IISAN-Versa variant of
Code_Cached_Asym/model/model.py:257-429 is selected with ``asym=True`` / the iisan_b200.model_asym package).

Everything numeric runs in libiisan_b200.so through the autograd functions of iisan_b200.ops.
"""
from __future__ import annotations

import torch
from torch import nn

from ..ops import InBatchCeFn, SanFn
from ..plan import SanBinder, make_plan
from ..precision import compute_mode
from .encoders import MM_Encoder, User_Encoder
from .modules import AdapterBlock, FusedLinear

ASYM_DEFAULT = False


class IISANAdaptedMModel(nn.Module):
    """Decoupled side-adapter network over cached per-layer hidden states.

    forward(sample_items_images, sample_items_text) -> (cv [N,E], [text [N,E], mm [N,E]]) like the
    reference; ``embed`` returns the same three blocks as one [N, 3E] tensor (the operand of com_dense).
    """

    def __init__(self, mm_model, args, asym=None):
        super().__init__()
        asym = ASYM_DEFAULT if asym is None else asym
        self.args = args
        plan = make_plan(args, asym)
        self.plan = plan
        E = plan.emb
        if asym:
            # CA/model/model.py:263-264: fresh E->E projections
            self.cv_pre_fc = nn.Linear(E, E)
            self.bert_pre_fc = nn.Linear(E, E)
        else:
            # CC/model/model.py:261-262: take ownership of the backbone heads (Linear 768->E)
            self.cv_pre_fc = mm_model.cv_encoder.image_net.classifier
            self.bert_pre_fc = mm_model.bert_encoder.text_encoders.title.fc
        self.side_bert_adapter_num_list = list(plan.layers_text_sel)
        self.side_cv_adapter_num_list = list(plan.layers_img_sel)
        drop = args.adapter_dropout_rate
        self.cv_adapter_list = nn.ModuleList(AdapterBlock(args, plan.d_img, plan.r_img, drop) for _ in range(plan.n_img))
        self.bert_adapter_list = nn.ModuleList(AdapterBlock(args, plan.d_text, plan.r_text, drop) for _ in range(plan.n_text))
        if plan.n_down_project:
            wide = max(plan.d_text, plan.d_img)
            self.down_project_list = nn.ModuleList(nn.Linear(wide, plan.d_mm) for _ in range(plan.n_down_project))
        self.mm_adapter_list = nn.ModuleList(AdapterBlock(args, plan.d_mm, plan.r_mm, drop) for _ in range(plan.n_mm))
        if asym:
            self.fc_bert = nn.Linear(plan.d_text, E)
            self.fc_cv = nn.Linear(plan.d_img, E)
        else:
            self.fc_bert = nn.Linear(plan.d_text, plan.d_text)
            self.fc_cv = nn.Linear(plan.d_img, plan.d_img)
        self.fc_mm = nn.Linear(plan.d_mm, plan.d_mm)
        self.fc_mm_down = nn.Linear(plan.d_mm, E)
        zero_gate = lambda: nn.Parameter(torch.zeros(1))           # sigmoid(0/0.1) = 0.5 at init
        self.side_gate_params_text = nn.ParameterList(zero_gate() for _ in range(plan.n_gate_text))
        self.side_gate_params_cv = nn.ParameterList(zero_gate() for _ in range(plan.n_gate_img))
        self.side_gate_params_mm = nn.ParameterList(zero_gate() for _ in range(plan.n_gate_mm))
        self._binder = None

    def _bind(self):
        if self._binder is None:
            self._binder = SanBinder(self.plan, [n for n, _ in self.named_parameters()])
        return self._binder

    def embed(self, sample_items_images, sample_items_text, packed=False):
        """``packed``: the tensors hold only the selected layers [N, A, d] (iisan_b200.store.CachedStateStore.gather)."""
        params = tuple(self.parameters())
        return SanFn.apply(self._bind(), sample_items_images, sample_items_text, compute_mode(), bool(packed), *params)

    def forward(self, sample_items_images, sample_items_text):
        out = self.embed(sample_items_images, sample_items_text)
        E = self.plan.emb
        return out[:, :E], [out[:, E:2 * E], out[:, 2 * E:]]


class ModelMM(nn.Module):
    """Code_Cached/model/model.py:15-105: item embedding fusion, SASRec user encoder and in-batch CE.

    Extra, reference-compatible knobs (all default to the reference behaviour):
      ``negatives``: "local" (per-rank in-batch negatives, the reference under DDP) or "global"
      (all-gathered item pool, see iisan_b200.parallel).
    """

    def __init__(self, args, item_num, use_modal, image_net, bert_model, pop_prob_list):
        super().__init__()
        self.args = args
        self.use_modal = use_modal
        self.max_seq_len = args.max_seq_len
        self.l2_weight = args.l2_weight / 2
        self.pop_prob_list = torch.as_tensor(pop_prob_list, dtype=torch.float32)
        self.user_encoder = User_Encoder(item_num=item_num, max_seq_len=args.max_seq_len, item_dim=args.embedding_dim,
                                         num_attention_heads=args.num_attention_heads, dropout=args.drop_rate,
                                         n_layers=args.transformer_block)
        if not use_modal:
            raise NotImplementedError("the id-embedding recommender (use_modal=False) is outside the IISAN(Cached) hot path")
        self.mm_encoder = MM_Encoder(args, image_net, bert_model)
        E = args.embedding_dim
        if "intra_inter" in args.modality:
            width = 3 * E
        elif "inter" in args.modality:
            width = E
        else:
            width = 2 * E
        self.com_dense = FusedLinear(width, E)
        self.criterion = nn.CrossEntropyLoss()       # kept for state/attribute parity; the fused loss replaces it
        self.negatives = "local"
        self.process_group = None

    def _pop(self, device):
        if self.pop_prob_list.device != device:
            self.pop_prob_list = self.pop_prob_list.to(device)
        return self.pop_prob_list

    def item_embeddings(self, sample_items_images, sample_items_text, packed=False):
        """score_embs [N, E] = com_dense(cat[cv, text, mm])  (model.py:66-72)."""
        enc = self.mm_encoder
        if not isinstance(enc, IISANAdaptedMModel):
            raise NotImplementedError("install the side-adapter network first: model.mm_encoder = "
                                      "IISANAdaptedMModel(model.mm_encoder, args)  (Code_Cached/run.py:182-183)")
        return self.com_dense(enc.embed(sample_items_images, sample_items_text, packed))

    def forward(self, sample_items_id, sample_items_images, sample_items_text, log_mask, local_rank=None, packed=False):
        E, S = self.args.embedding_dim, self.max_seq_len + 1
        score_embs = self.item_embeddings(sample_items_images, sample_items_text, packed)
        device = score_embs.device
        ids = sample_items_id.to(device).view(-1)
        log_mask = log_mask.to(device=device, dtype=torch.float32)
        pool = deferred = None
        if self.negatives == "global":
            # start the pool all-gather now: the SASRec forward below only needs the local rows and hides the transfer; in the
            # backward the reduce-scatter of d score_all is launched before the SASRec backward and joined after it
            from .. import parallel as par
            import torch.distributed as dist
            group = self.process_group if self.process_group is not None else dist.group.WORLD
            if score_embs.dtype == torch.float32 and dist.get_backend(group) != "gloo":
                deferred = par._PendingScatter()
                if score_embs.requires_grad:
                    score_embs = par.JoinPoolGrad.apply(score_embs, deferred)
                else:
                    deferred = None
                pool = par.start_pool_gather(score_embs, ids.view(log_mask.shape[0], -1), log_mask, group)
        input_embs = score_embs.view(-1, S, E)
        prec_vec = self.user_encoder(input_embs, log_mask, local_rank, seq_len=S - 1).reshape(-1, E)   # model.py:76-79: input_embs[:, :-1, :]
        pop = self._pop(device)
        if self.negatives == "global":
            from ..parallel import global_negative_loss
            return global_negative_loss(prec_vec, score_embs, ids, log_mask, pop, self.process_group, compute_mode(), pool=pool,
                                        deferred=deferred)
        _sum, _n, loss = InBatchCeFn.apply(prec_vec, score_embs, ids, ids, log_mask, log_mask, pop, 0, compute_mode())
        return loss


class Model(nn.Module):
    """Placeholder for `from model import *` parity: the uncached model is out of scope."""

    def __init__(self, *a, **k):
        raise NotImplementedError("Model (uncached backbones) is outside the IISAN(Cached) hot path; use ModelMM")
