"""ORACLE - test infrastructure, build-container only.

Imports the reference's own modules from /root/reference/backend so that the
oracle restatement can be pinned against them and golden vectors generated.
/root/reference does not exist on the GPU box: nothing that runs there imports
this file.

The reference cannot be imported unmodified offline:
  * `import hopsworks`, `hsml`, `boto3` at module top (training_pipeline.py:3,7-9,37,67)
    -> empty stub modules;
  * `AutoTokenizer.from_pretrained("bert-base-uncased")` at import (:323)
    -> the seeded local vocabulary of mmdx_b200.synth;
  * constructors that download weights (`tv.resnet50(weights=...)` :178,
    `AutoConfig/AutoModel.from_pretrained` :358-360, `T5ForConditionalGeneration
    .from_pretrained` :545) -> same architectures, random init, no download.
No reference source is copied; the modules are executed where they lie.
"""
import sys
import types

REFERENCE_BACKEND = "/root/reference/backend"
_cached = None


def import_reference():
    """Returns (training_pipeline, inference_pipeline) modules of the reference."""
    global _cached
    if _cached is not None:
        return _cached
    import torchvision.models as tvm
    import transformers
    from transformers import BertConfig, BertModel, T5Config, T5ForConditionalGeneration

    from mmdx_b200 import synth

    for name in ("hopsworks", "hsml", "hsml.schema", "hsml.model_schema", "boto3"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["hsml.schema"].Schema = object
    sys.modules["hsml.model_schema"].ModelSchema = object
    sys.modules["boto3"].client = lambda *a, **k: None

    tok = synth.make_bert_tokenizer()
    transformers.AutoTokenizer.from_pretrained = classmethod(lambda cls, *a, **k: tok)
    transformers.AutoConfig.from_pretrained = classmethod(lambda cls, *a, **k: BertConfig())
    transformers.AutoModel.from_pretrained = classmethod(lambda cls, *a, **k: BertModel(BertConfig()))
    t5cfg = T5Config(decoder_start_token_id=0)
    T5Config.from_pretrained = classmethod(lambda cls, *a, **k: t5cfg)
    T5ForConditionalGeneration.from_pretrained = classmethod(lambda cls, *a, **k: T5ForConditionalGeneration(t5cfg))
    orig_resnet50 = tvm.resnet50
    tvm.resnet50 = lambda weights=None, **k: orig_resnet50(weights=None, **k)

    if REFERENCE_BACKEND not in sys.path:
        sys.path.insert(0, REFERENCE_BACKEND)
    from ml.pipelines import inference_pipeline, training_pipeline
    _cached = (training_pipeline, inference_pipeline)
    return _cached


def build_reference_bundle(state_bundle: dict) -> dict:
    """Instantiate the reference's three nn.Modules and load the seeded state dicts into
    them the way api/views.py:216-234 does; returns the serving bundle inference() reads."""
    tp, _ = import_reference()
    fc = state_bundle["cfg"]["fusion"]
    img = tp.ImageEncoderCNN(d_img=fc["d_img"], n_disease_classes=fc["n_disease"])
    img.load_state_dict(state_bundle["image_state"], strict=True)
    txt = tp.TextEncoderTransformer(d_txt=fc["d_txt"], n_disease=fc["n_disease"])
    missing, unexpected = txt.load_state_dict(state_bundle["text_state"], strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    fus = tp.FusionTransformerModel(d_img=fc["d_img"], d_txt=fc["d_txt"], d_fuse_hidden=fc["d_fuse_hidden"],
                                    n_disease=fc["n_disease"], init_t5_from_config=True)
    missing, unexpected = fus.load_state_dict(state_bundle["fusion_state"], strict=False)
    assert not unexpected and all(m.startswith("report_model.") for m in missing), (missing, unexpected)
    return {"fusion_model": fus.eval(), "image_encoder": img.eval(), "text_encoder": txt.eval(),
            "t5_tok": None, "bert_tok": tp.tokenizer, "class_names": state_bundle["class_names"],
            "thresholds": state_bundle["thresholds"], "version": state_bundle["version"], "cfg": state_bundle["cfg"]}
