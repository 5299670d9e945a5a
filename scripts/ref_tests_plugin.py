"""pytest plugin used by scripts/run_reference_tests.py: environment for running the reference's OWN, unmodified
test files on top of this repository's drop-in `quant` package on a GPU box.

The reference's tests build CPU tensors (`torch.ones(...)`) and the task tests ask for `ngpus: 0`; the B200
implementation has no CPU path by design, so the plugin only changes WHERE things live -- never what is asserted:
  * the default torch device becomes cuda:0, so tensors and modules the tests construct are CUDA tensors;
  * `tests.data.helpers.get_base_config_template` keeps returning the reference's config with
    environment.ngpus = 1 (one GPU instead of the CPU).
Known, intended failures: tests/utils/test_moving_average.py::test_moving_average_{train_and_eval,eval_only} loop over
an explicit torch.device('cpu') first; CPU tensors are rejected by design (no CPU fallback), so they stop there.  Their
CUDA half is covered by tests/test_gpu_round2.py::test_moving_average_reference_closed_form_on_cuda.
"""
import torch


def pytest_configure(config):
    assert torch.cuda.is_available(), 'the reference tests run against the CUDA implementation: a GPU is required'
    torch.set_default_device('cuda:0')
    # tests/binary/test_binary_conv.py::test_fp_quant_conv2d_eq_nn_conv2d compares the input gradients of two plain
    # F.conv2d calls bit for bit; cuDNN's default dgrad algorithm is not run-to-run deterministic on a GPU
    torch.backends.cudnn.deterministic = True
    import quant.binary.binary_conv as bc
    assert 'ml_quant_b200' in bc.QuantConv2d.__module__, bc.QuantConv2d.__module__    # the shim, not the reference
    import tests.data.helpers as helpers
    orig = helpers.get_base_config_template

    def on_one_gpu(*a, **k):
        cfg = orig(*a, **k)
        cfg['environment']['ngpus'] = 1
        return cfg
    helpers.get_base_config_template = on_one_gpu
