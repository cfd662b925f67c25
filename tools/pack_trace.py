"""Which modules still convert NCHW tensors into the operand format (pack_nchw / im2col / spade_pack launches of one pass):
    python tools/pack_trace.py [batch]"""
import collections, importlib, inspect, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda', 0)
G = bench.build_generator(dev)
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
plugin = custom_ops.get_plugin('conv2d_plugin')
log = []


def owner():
    names = []
    for fr in inspect.stack()[2:]:
        slf = fr.frame.f_locals.get('self')
        if isinstance(slf, torch.nn.Module):
            names.append(f'{type(slf).__name__}.{fr.function}')
        if len(names) == 3:
            break
    return ' < '.join(names)


def wrap(name):
    orig = getattr(plugin, name)

    def f(x, *a, **k):
        log.append((name, tuple(x.shape), owner()))
        return orig(x, *a, **k)
    setattr(plugin, name, staticmethod(f))


for nm in ('pack_activations', 'pack_activations_into', 'pack_im2col', 'spade_modulate_pack'):
    wrap(nm)
x = bench.to_device_f32(bench.make_generator_inputs_u8(batch, 100), dev)
bench.run_generator(G, x)
torch.cuda.synchronize()
tot = collections.Counter()
for name, shape, who in log:
    print(f'{name:22s} {str(shape):28s} {who}')
    tot[name] += shape[0] * shape[1] * shape[2] * shape[3]
print({k: f'{v * 4 / 1e6 / batch:.1f} MB/img fp32 in' for k, v in tot.items()})
