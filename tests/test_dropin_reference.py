"""CPU, only where the reference checkout exists (this container, not the GPU box): the drop-in modules really are
subclasses of the reference's own CRModule -- same constructor keywords, same parameter names (so Lightning checkpoints
load), B200 hooks first in the method resolution order.  The reference's heavy dependencies are shimmed exactly as for the
golden generator (oracle/ref_stubs.py); the PLM encoder is replaced by a small module.  Runs in a subprocess: installing the
shims changes sys.modules / sys.path for good."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "manner")), reason="reference checkout not present")

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import torch
from oracle import ref_stubs
ref_stubs.install(%(ref)r)
import manner.models.cr_module as ref_cr


class TinyEncoder(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.proj = torch.nn.Linear(4, 8)


ref_cr.MannerNewsEncoder = TinyEncoder
from manner_b200 import modules
assert modules.HAVE_REFERENCE
KW = dict(supcon_loss=True, late_fusion=False, temperature=0.36, plm_model="", frozen_layers=[], dropout_probability=0.2,
          use_entities=False, pretrained_entity_embeddings_path="", entity_embedding_dim=100, num_attention_heads=10,
          query_vector_dim=16, text_embedding_dim=8, optimizer=None)
ref = ref_cr.CRModule(**KW)
ours = modules.CRModuleB200(**KW)  # same keywords; `scorer` is defaulted
assert isinstance(ours, ref_cr.CRModule)
assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())  # checkpoints of the reference load unchanged
ours.load_state_dict(ref.state_dict())
assert ours.hparams["temperature"] == 0.36 and ours.hparams["late_fusion"] is False  # the reference's own hparams are recorded
mro = [c.__name__ for c in type(ours).__mro__]
assert mro.index("B200EvalMixin") < mro.index("CRModule")
assert ours._b200_loss() == "supcon" and ours._b200_cached is False
att = ours._b200_attention()  # early fusion: the module's own additive-attention parameters
assert att is not None and att[0][0] is ours.user_encoder.additive_attention.linear.weight
assert att[0][2] is ours.user_encoder.additive_attention.query
late = modules.CRModuleB200(**dict(KW, late_fusion=True, supcon_loss=False), scorer="b200_cached")
assert late._b200_attention() is None and late._b200_loss() == "ce" and late._b200_cached is True
assert modules.CRModuleB200(**KW, scorer="reference")._b200_enabled is False
try:
    modules.CRModuleB200(**KW, scorer="cpu")
    raise SystemExit("a bad scorer was accepted")
except ValueError:
    pass

# ---- EnsembleModuleB200 on the reference's EnsembleModule (sub-modules come from checkpoints: patched like make_golden.py) ----
import manner.models.a_module as ref_a
import manner.models.ensemble_module as ref_ens


class Sub(torch.nn.Module):
    def __init__(self, tag):
        super().__init__()
        self.news_encoder = TinyEncoder()
        self.tag = tag


ref_cr.CRModule.load_from_checkpoint = classmethod(lambda cls, checkpoint_path, **kw: Sub(checkpoint_path))
ref_a.AModule.load_from_checkpoint = classmethod(lambda cls, checkpoint_path, **kw: Sub(checkpoint_path))
EKW = dict(cr_module_module_ckpt="cr", a_module_categ_ckpt="categ", a_module_sent_ckpt="sent", categ_weight=0.3, sent_weight=0,
           num_categ_classes=19, num_sent_classes=4)
ens = modules.EnsembleModuleB200(**EKW)
assert isinstance(ens, ref_ens.EnsembleModule) and ens.hparams["categ_weight"] == 0.3
assert list(ens.state_dict().keys()) == list(ref_ens.EnsembleModule(**EKW).state_dict().keys())
encs = ens._b200_encoders()
assert len(encs) == 2 and encs[0] is ens.cr_module.news_encoder and encs[1] is ens.a_module_categ.news_encoder  # sent_weight 0: not loaded
assert ens._b200_weights() == [1.0, 0.3] and ens._b200_zscore and not ens._b200_with_auc and ens._b200_loss() is None
assert modules.EnsembleModuleB200(**dict(EKW, sent_weight=0.5), scorer="b200_cached")._b200_weights() == [1.0, 0.3, 0.5]
# ---- model.aspect_weights: a sweep through the reference's real EnsembleModule constructor --------------------------------
sweep = modules.EnsembleModuleB200(**dict(EKW, categ_weight=0, sent_weight=0), aspect_weights=[[0, 0], [0.2, 0], [0.4, 0.1]])
assert isinstance(sweep, ref_ens.EnsembleModule) and sweep.hparams["categ_weight"] == 0  # the reference's own hparams are untouched
assert hasattr(sweep, "a_module_categ") and hasattr(sweep, "a_module_sent")  # loaded for the sweep although the scalar weights are 0
assert sweep.a_module_categ.tag == "categ" and sweep.a_module_sent.tag == "sent"
assert sweep._b200_weights() == [[1.0, 0.0, 0.0], [1.0, 0.2, 0.0], [1.0, 0.4, 0.1]] and len(sweep._b200_encoders()) == 3
only_c = modules.EnsembleModuleB200(**dict(EKW, categ_weight=0, sent_weight=0), aspect_weights=[[0.1, 0], [0.3, 0]])
assert not hasattr(only_c, "a_module_sent") and only_c._b200_weights() == [[1.0, 0.1], [1.0, 0.3]]
try:
    modules.EnsembleModuleB200(**EKW, aspect_weights=[[0.1, 0.2, 0.3]])
    raise SystemExit("a malformed aspect_weights was accepted")
except ValueError:
    pass

# ---- B200MetricsMixin on a real baseline of the reference (nrms_plm_module.py) ---------------------------------------------
import manner.models.baselines.nrms_plm_module as ref_nrms

ref_nrms.NewsEncoder = TinyEncoder
ref_nrms.UserEncoder = TinyEncoder


class NRMSModuleB200(modules.B200MetricsMixin, ref_nrms.NRMSPLMModule):
    pass


nrms = NRMSModuleB200(plm_model="", frozen_layers=[], dropout_probability=0.2, text_embedding_dim=16, num_attention_heads=2, query_vector_dim=8,
                      num_categ_classes=19, num_sent_classes=4, optimizer=None)
mro = [c.__name__ for c in type(nrms).__mro__]
assert mro.index("B200MetricsMixin") < mro.index("NRMSPLMModule")
assert type(nrms).on_test_epoch_end is modules.B200MetricsMixin.on_test_epoch_end and type(nrms).test_step is ref_nrms.NRMSPLMModule.test_step
assert set(nrms.keys) == set(nrms.test_step_outputs) and "hist_sentiments" in nrms.keys  # the buffers the mixin reads
nrms.test_step_outputs["preds"].append(torch.zeros(3))
try:
    for k in nrms.keys[1:]:
        nrms.test_step_outputs[k].append(torch.zeros(3, dtype=torch.long))
    nrms.on_test_epoch_end()
    raise SystemExit("the mixin computed on CPU tensors")
except RuntimeError as e:
    assert "no CPU path" in str(e)
print("dropin ok")
'''


def test_cr_module_b200_is_the_reference_module_with_b200_hooks():
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "ref": REFERENCE}], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "dropin ok" in out.stdout, out.stderr[-3000:]
