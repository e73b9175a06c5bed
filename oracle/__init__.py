"""CPU oracle for the ViT / DeiT / T2T inference forward.

TEST INFRASTRUCTURE ONLY.  Nothing in ``edgevisiontransformer_b200`` (the
product) may import this package; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and
there only as the checker or as the timed CPU baseline.

What it restates (citations are paths under /root/reference, or ``SITE`` =
the installed ``transformers`` package whose arithmetic the reference calls):

* ``oracle.vit``      HF ``ViTForImageClassification`` forward
                      (``SITE/models/vit/modeling_vit.py:100-128,151-167,171-196,
                      220-251,265-268,296-312,328-346,455,620-653``), third-party
                      dependency pinned ``transformers==4.7.0`` at
                      ``deit_pruning/requirements.txt:21``; call sites
                      ``deit_pruning/src/utils.py:194-195``,
                      ``are_16_heads/classifier_eval.py:69-70``.
* ``oracle.pruning``  HF-4.x ``prune_heads`` and nn_pruning ``optimize_model``
                      (``deit_pruning/vendor/nn_pruning_v1/nn_pruning/
                      inference_model_patcher.py:22-89,266-317``).
* ``oracle.torch_layers`` ``modeling/torch_layers/*`` as composed by
                      ``utils.py:322-365``.
* ``oracle.tf_vit``   TF-dialect DeiT (``modeling/models/vit.py:9-109``,
                      ``modeling/layers/*``).
* ``oracle.t2t``      T2T-ViT (``modeling/models/t2t_vit.py:7-148``,
                      ``modeling/layers/transformer_encoder.py:39-101``).
* ``oracle.timm_vit`` timm / facebookresearch-deit ``VisionTransformer`` (the model
                      ``utils.py:52-62`` loads over torch.hub; not under /root/reference).
* ``oracle.swin``     Swin Transformer classifier (the model ``utils.py:14-47`` builds from
                      an external microsoft/Swin-Transformer checkout; arithmetic as in
                      ``SITE/models/swin/modeling_swin.py``).

Pinning status
--------------
The reference holds NO golden vector / known-answer test for this path
(SURVEY.md section 4).  Pins are therefore outputs of the reference itself run in
the build container, committed under ``tests/golden/`` with the generating
script ``tests/golden/make_golden.py``:

* ``oracle.vit`` / ``oracle.pruning``  pinned against the installed
  ``transformers`` ViT forward and the vendored ``optimize_model`` imported
  from /root/reference (fixtures ``hf_*.npz``, ``pruned_*.npz``).
* ``oracle.torch_layers``  pinned against ``/root/reference/modeling/torch_layers``
  imported unchanged (fixtures ``torch_layers_*.npz``).
* ``oracle.timm_vit`` / ``oracle.swin``  pinned against the installed ``transformers``
  implementations of the same networks (fixtures ``timm_tiny_s5.npz``, ``swin_tiny_*.npz``);
  unpinned against timm / the microsoft repository themselves (external, not fetchable).
* ``oracle.tf_vit`` / ``oracle.t2t``  **parity unpinned**: TensorFlow is not
  installed, the reference's implementation cannot run here, and the
  reference ships no fixture for it.  They are line-by-line restatements only.
"""

from .spec import ViTSpec  # noqa: F401
