"""Checkpoint loading: `network-snapshot-*.pkl` of the reference -> this package's generator.

Mirror of the reference's `legacy.load_network_pkl` (legacy.py:20-60, used by test.py:104-105) with one deliberate
difference: the reference unpickles by EXECUTING the source code embedded in the pickle (`torch_utils/persistence.py:179-227`,
`_reconstruct_persistent_obj` -> `_src_to_module` -> `exec`).  Here nothing from the file is executed or imported: a
restricted unpickler maps every persistent object to an inert `PersistentRecord` (class name, embedded source text, init
args, module state), tensors are rebuilt through torch's own weights-only loader, and any global outside a small
allow-list raises `pickle.UnpicklingError`.  The records are then turned into this package's modules by class name.

    data = load_network_pkl(f)                      # {'G': record, 'D': record, 'G_ema': record, 'training_set_kwargs': ..., ...}
    G = build_generator(data['G_ema'])              # GeneratorFull_v20 of this package with the snapshot's weights
    src = class_source(data['G_ema'], 'SynthesisLayer')   # text of a class from the embedded module source (never executed)
"""
import ast
import collections
import inspect
import io
import pickle

import numpy as np
try:
    from numpy import _core as _np_core
    _np_multiarray = _np_core.multiarray
except ImportError:      # numpy < 2
    _np_multiarray = np.core.multiarray
import torch


class EasyDict(dict):
    """attribute-access dict (stands in for dnnlib.util.EasyDict found in snapshots)"""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


class PersistentRecord:
    """One object saved through the reference's `persistence.persistent_class` decorator (persistence.py:118-126)."""

    def __init__(self, meta):
        assert meta.get('type') == 'class', f'unsupported persistent object type {meta.get("type")!r}'
        self.class_name = meta['class_name']
        self.module_src = meta['module_src']
        self.version = meta.get('version')
        self.state = dict(meta['state'])

    @property
    def init_args(self):
        return tuple(self.state.get('_init_args', ()))

    @property
    def init_kwargs(self):
        return EasyDict(self.state.get('_init_kwargs', {}))

    def __repr__(self):
        return f'PersistentRecord({self.class_name}, {len(state_dict(self))} tensors)'


class PlainModuleRecord:
    """A torch.nn module (or a non-persistent class of training.networks) pickled by value: only its state is kept."""
    class_path = None

    def __setstate__(self, state):
        self.state = dict(state)

    @property
    def class_name(self):
        return self.class_path.rsplit('.', 1)[-1]


_plain_records = {}


def _plain_record_class(module, name):
    key = f'{module}.{name}'
    if key not in _plain_records:
        _plain_records[key] = type(name, (PlainModuleRecord,), dict(class_path=key))
    return _plain_records[key]


def _load_storage_from_bytes(b):
    # torch.storage._load_from_bytes would call torch.load(weights_only=False); the blob only needs the restricted loader
    return torch.load(io.BytesIO(b), weights_only=True)


_ALLOWED = {
    ('collections', 'OrderedDict'): collections.OrderedDict,
    ('torch._utils', '_rebuild_tensor_v2'): torch._utils._rebuild_tensor_v2,
    ('torch._utils', '_rebuild_parameter'): torch._utils._rebuild_parameter,
    ('torch._utils', '_rebuild_parameter_with_state'): torch._utils._rebuild_parameter_with_state,
    ('torch.storage', '_load_from_bytes'): _load_storage_from_bytes,
    ('torch', 'Size'): torch.Size,
    ('torch', 'device'): torch.device,
    ('torch_utils.persistence', '_reconstruct_persistent_obj'): lambda meta: PersistentRecord(meta),
    ('dnnlib.util', 'EasyDict'): EasyDict,
    ('numpy.core.multiarray', '_reconstruct'): _np_multiarray._reconstruct,
    ('numpy._core.multiarray', '_reconstruct'): _np_multiarray._reconstruct,
    ('numpy.core.multiarray', 'scalar'): _np_multiarray.scalar,
    ('numpy._core.multiarray', 'scalar'): _np_multiarray.scalar,
    ('numpy', 'ndarray'): np.ndarray,
    ('numpy', 'dtype'): np.dtype,
    ('copyreg', '_reconstructor'): lambda cls, base, state: cls.__new__(cls),
    ('copy_reg', '_reconstructor'): lambda cls, base, state: cls.__new__(cls),
    ('builtins', 'object'): object,
    ('__builtin__', 'object'): object,
    ('builtins', 'set'): set,
    ('builtins', 'frozenset'): frozenset,
}
_TORCH_DTYPES = ('float32', 'float64', 'float16', 'bfloat16', 'int64', 'int32', 'int16', 'int8', 'uint8', 'bool')
_TORCH_STORAGES = ('FloatStorage', 'DoubleStorage', 'HalfStorage', 'BFloat16Storage', 'LongStorage', 'IntStorage', 'ShortStorage',
                   'CharStorage', 'ByteStorage', 'BoolStorage')


class _RestrictedUnpickler(pickle.Unpickler):
    plain_modules = ('training.networks', 'training.augment')

    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            return _ALLOWED[(module, name)]
        if module == 'torch' and (name in _TORCH_DTYPES or name in _TORCH_STORAGES):
            return getattr(torch, name)
        if (module.startswith('torch.nn.modules.') or module in self.plain_modules) and name.isidentifier():
            return _plain_record_class(module, name)        # inert stand-in: state only, no code from the module runs
        raise pickle.UnpicklingError(f'snapshot refers to {module}.{name}, which is not on the allow-list of this loader '
                                     '(nothing in a snapshot is imported or executed)')


def load_network_pkl(f, force_fp16=False, plain_modules=()):
    """Same call as the reference's `legacy.load_network_pkl(f)`; the networks come back as inert records (see module
    docstring) - pass them to `build_generator`.  `force_fp16` is accepted for signature parity and must be False: this
    package's generator keeps fp32 semantics and picks its tensor-core precision through `conv2d_gradfix.fp32_precision`.
    `plain_modules`: extra module names whose (non-persistent) classes may appear in the file as state-only records."""
    assert not force_fp16, 'force_fp16 is not supported: precision is selected by conv2d_gradfix.fp32_precision'
    unpickler = _RestrictedUnpickler(f)
    unpickler.plain_modules = _RestrictedUnpickler.plain_modules + tuple(plain_modules)
    data = unpickler.load()
    if not isinstance(data, dict):
        raise pickle.UnpicklingError('not a PASTA-GAN++ / StyleGAN2-ADA PyTorch snapshot (TensorFlow-era pickles are not supported)')
    data.setdefault('training_set_kwargs', None)
    data.setdefault('augment_pipe', None)
    for key in ('G', 'D', 'G_ema'):
        if key in data and not isinstance(data[key], PersistentRecord):
            raise pickle.UnpicklingError(f'snapshot entry {key!r} is not a persistent network')
    return data


def state_dict(record, prefix=''):
    """Flat `name -> tensor` dict of a record, with torch.nn.Module.state_dict() naming (persistent buffers only)."""
    out = collections.OrderedDict()
    st = record.state
    for name, p in (st.get('_parameters') or {}).items():
        if p is not None:
            out[prefix + name] = p.data if isinstance(p, torch.nn.Parameter) else p
    skip = set(st.get('_non_persistent_buffers_set') or ())
    for name, b in (st.get('_buffers') or {}).items():
        if b is not None and name not in skip:
            out[prefix + name] = b
    for name, child in (st.get('_modules') or {}).items():
        if child is not None:
            out.update(state_dict(child, prefix + name + '.'))
    return out


def module_tree(record, prefix=''):
    """`path -> class name` of every sub-module of a record (for inspection and for checking a snapshot's architecture)."""
    out = collections.OrderedDict([(prefix.rstrip('.'), record.class_name)])
    for name, child in (record.state.get('_modules') or {}).items():
        if child is not None:
            out.update(module_tree(child, prefix + name + '.'))
    return out


def class_source(record, class_name):
    """Source text of `class_name` from the module source embedded in the snapshot (parsed with `ast`, never executed).
    This is how the upstream `SynthesisLayer` - absent from the reference tree - can be read out of a released snapshot."""
    tree = ast.parse(record.module_src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            return ast.get_source_segment(record.module_src, node)
    raise KeyError(f'{class_name} is not defined in the embedded source of {record.class_name}')


def default_registry():
    """explicit allow-list of the network classes a snapshot may name (never helper functions or other module globals)"""
    from .training import augment, discriminator, generator, synthesis
    names = {generator: ('GeneratorFull_v20', 'SynthesisNetworkFull_v18', 'SynthesisBlockFull', 'MappingNetwork', 'ResBlock',
                         'ConstEncoderNetwork', 'StyleEncoderNetworkV18', 'Dense', 'Spade_Conv2dLayer', 'Spade_Norm_Block',
                         'Spade_ResBlockV4_512'),
             synthesis: ('FullyConnectedLayer', 'Conv2dLayer', 'SynthesisLayer', 'ToRGBLayer', 'SynthesisBlock', 'SynthesisChain'),
             discriminator: ('Discriminator', 'DiscriminatorBlock', 'DiscriminatorEpilogue', 'MinibatchStdLayer'),
             augment: ('AugmentPipe',)}
    return {n: getattr(mod, n) for mod, ns in names.items() for n in ns}


def build_module(record, registry=None):
    """Instantiate this package's class of the same name with the record's init arguments and load its weights."""
    if registry is None:
        registry = default_registry()
    cls = registry.get(record.class_name)
    if cls is None or not (isinstance(cls, type) and issubclass(cls, torch.nn.Module)):
        raise KeyError(f'this package has no network class named {record.class_name}')
    # class name and arguments come from the (untrusted) file: only torch.nn.Module classes on the allow-list are
    # instantiated, and only with arguments their constructor declares
    try:
        inspect.signature(cls.__init__).bind(None, *record.init_args, **record.init_kwargs)
    except TypeError as e:
        raise pickle.UnpicklingError(f'snapshot arguments do not match {record.class_name}.__init__: {e}') from None
    module = cls(*record.init_args, **record.init_kwargs)
    module.load_state_dict(state_dict(record), strict=True)
    return module.eval().requires_grad_(False)


def build_generator(record):
    """GeneratorFull_v20 of this package from the `G_ema` (or `G`) record of a snapshot (test.py:104-105)."""
    if record.class_name != 'GeneratorFull_v20':
        raise ValueError(f'expected a GeneratorFull_v20 snapshot, found {record.class_name}')
    return build_module(record)
