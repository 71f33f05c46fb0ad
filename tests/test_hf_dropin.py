"""Drop-in check one level above the module: Hugging Face's own `replace_with_aqlm_linear`
(transformers/integrations/aqlm.py:26-70) does `from aqlm import QuantizedLinear` and constructs it with keywords on the
meta device.  With `aqlm_b200.install_as_aqlm()` it must build OUR module, and a state_dict in the reference's format
(convert_to_hf.py:59-68: names codes / codebooks / scales, packed int codes, fp16 floats) must load by name."""
import sys

import pytest
import torch

transformers = pytest.importorskip("transformers")


@pytest.fixture
def aqlm_alias():
    import aqlm_b200

    saved = {k: v for k, v in sys.modules.items() if k == "aqlm" or k.startswith("aqlm.")}
    aqlm_b200.install_as_aqlm()
    yield aqlm_b200
    for k in [k for k in sys.modules if k == "aqlm" or k.startswith("aqlm.")]:
        del sys.modules[k]
    sys.modules.update(saved)


def test_replace_with_aqlm_linear_builds_our_module(aqlm_alias):
    from transformers import AqlmConfig, LlamaConfig, LlamaForCausalLM
    from transformers.integrations.aqlm import replace_with_aqlm_linear

    cfg = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=64)
    with torch.device("meta"):
        model = LlamaForCausalLM(cfg)
    qcfg = AqlmConfig(in_group_size=8, out_group_size=1, num_codebooks=1, nbits_per_codebook=16)
    model = replace_with_aqlm_linear(model, modules_to_not_convert=["lm_head"], quantization_config=qcfg)
    layer = model.model.layers[0]
    for name, fin, fout in [("self_attn.q_proj", 128, 128), ("self_attn.k_proj", 128, 64), ("mlp.gate_proj", 128, 256),
                            ("mlp.down_proj", 256, 128)]:
        m = layer.get_submodule(name)
        assert type(m) is aqlm_alias.QuantizedLinear, (name, type(m))
        assert (m.in_features, m.out_features) == (fin, fout)
        assert m.codes.shape == (fout, fin // 8, 1) and m.codes.dtype == torch.int16 and m.codes.device.type == "meta"
        assert m.codebooks.shape == (1, 65536, 1, 8) and m.scales.shape == (fout, 1, 1, 1) and m.bias is None
        assert not any(p.requires_grad for p in m.parameters())
    assert isinstance(model.lm_head, torch.nn.Linear)


def test_reference_format_state_dict_loads_by_name(aqlm_alias):
    from aqlm_b200.utils import pack_int_data

    m = aqlm_alias.QuantizedLinear(64, 32, 8, 1, 2, 8, bias=True, dtype=torch.float16)
    sd = {  # what convert_to_hf.py writes: floats -> fp16, integer codes -> pack_int_data
        "codes": pack_int_data(torch.randint(0, 256, (32, 8, 2)), 8),
        "codebooks": torch.randn(2, 256, 1, 8).half(),
        "scales": torch.rand(32, 1, 1, 1).half(),
        "bias": torch.randn(32).half(),
    }
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert m.codes.dtype == torch.int8 and torch.equal(m.codes, sd["codes"])
