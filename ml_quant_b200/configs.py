"""``arch_config`` dictionaries of the reference's shipped example configs (the model section of
examples/{mnist,cifar100,imagenet}/*.yaml), so benchmarks and tests can build the same networks on a
box without the reference checkout.  Keys and values are the YAML's."""
import copy


def _stage(x_quant, w_quant, alpha, double_shortcut=True):
    cfg = {'x_quant': x_quant, 'w_quant': w_quant, 'double_shortcut': double_shortcut}
    cfg['clamp'] = {'kind': 'identity'} if alpha is None else {'kind': 'symmetric', 'alpha': alpha}
    return cfg


def _resnet18(x_quant, w_quant, alpha, nonlin, imagenet, classes):
    if imagenet:
        layer0 = {'n_in_channels': 64, 'kernel_size': 7, 'stride': 2, 'padding': 3, 'bias': False,
                  'maxpool': {'type': 'maxpool2d', 'kernel_size': 3, 'stride': 2, 'padding': 1}}
    else:
        layer0 = {'n_in_channels': 64, 'kernel_size': 3, 'stride': 1, 'padding': 1, 'bias': False,
                  'maxpool': {'type': 'identity'}}
    arch = {'moving_average_mode': 'off', 'moving_average_momentum': 0.99, 'block': 'xnor', 'layer0': layer0,
            'nonlins': [nonlin, nonlin], 'num_blocks': [2, 2, 2, 2], 'output_classes': classes}
    for i in range(1, 5):
        arch[f'layer{i}'] = _stage(x_quant, w_quant, alpha)
    return arch


ARCH = {
    # examples/imagenet/imagenet_ls1_weight_ls2_activation_kd.yaml  (north-star config)
    'imagenet_resnet18_ls1w_ls2a': _resnet18('ls-2', 'ls-1', 3, 'relu', True, 1000),
    # examples/imagenet/imagenet_ls1_kd.yaml (XNOR config)
    'imagenet_resnet18_ls1w_ls1a': _resnet18('ls-1', 'ls-1', 2, 'prelu', True, 1000),
    # examples/imagenet/imagenet_ls1_weight_lsT_activation_kd.yaml / ..._gf2_...
    'imagenet_resnet18_ls1w_lsTa': _resnet18('ls-T', 'ls-1', 2, 'prelu', True, 1000),
    'imagenet_resnet18_ls1w_gf2a': _resnet18('gf-2', 'ls-1', 3, 'prelu', True, 1000),
    # examples/cifar100/cifar100_ls1_weight_ls2_activation_kd.yaml
    'cifar100_resnet18_ls1w_ls2a': _resnet18('ls-2', 'ls-1', 2, 'relu', False, 100),
    # examples/mnist/mnist_ls1_weight_fp_activation.yaml
    'mnist_lenet5_ls1w_fpa': {'moving_average_mode': 'off', 'moving_average_momentum': 0.99, 'x_quant': 'fp',
                              'w_quant': 'ls-1', 'clamp': {'kind': 'identity'}, 'conv1_filters': 20,
                              'conv2_filters': 50, 'output_classes': 10},
}

INPUT = {'imagenet': (3, 224, 224), 'cifar100': (3, 32, 32), 'mnist': (1, 28, 28)}


def arch(name):
    return copy.deepcopy(ARCH[name])


def input_shape(name):
    return INPUT[name.split('_')[0]]
