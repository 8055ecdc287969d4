"""Dev-container only: the reference's OWN source files (imported unmodified from /root/reference under the
todd shim) vs the oracle restatement, bit for bit, and the committed golden files are reproducible."""
import pathlib

import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present')]
GOLDEN = pathlib.Path(__file__).parent / 'golden'


def test_restatement_equals_reference_source_and_golden_is_reproducible():
    from oracle import make_golden as M
    for name, args in M.CASES.items():
        rec = M.run_case(name, *args)          # asserts oracle == reference inside
        disk = torch.load(GOLDEN / f'{name}.pt', weights_only=False)
        for a, b in zip(rec['steps'], disk['steps']):
            assert torch.equal(a['quant'], b['quant']) and torch.equal(a['z'], b['z'])
            assert torch.equal(a['W_after'], b['W_after']) and torch.equal(a['loss'], b['loss'])
    for levels in ([8, 8, 5, 5, 5], [8, 8, 8, 5, 5, 5]):
        M.run_fsq(levels)


def test_plugin_force_registers_into_reference_registries():
    """The drop-in seam: after importing the plugin, the reference's registry builds OUR classes from a
    reference-style config (custom_imports mechanism, vq/train.py:36-37)."""
    ref = ref_loader.load()
    import importlib
    import vector_quantization_b200.plugin as plugin
    importlib.reload(plugin)
    assert plugin.registered_into_reference
    import vector_quantization_b200 as vqb
    cfg = ref.todd.Config(type='VQGANQuantizer', distance=dict(type='L2Distance'),
                          embedding=dict(type='torch_nn_modules_sparse_Embedding', num_embeddings=64, embedding_dim=8),
                          losses=dict(vqgan_loss=dict(type='VQGANLoss')))
    q = ref.VQITQuantizerRegistry.build(cfg)
    assert type(q) is vqb.VQGANQuantizer and type(q.distance) is vqb.L2Distance
    assert type(q._losses['vqgan_loss']) is vqb.VQGANLoss
