"""A light net IR standing in for caffe2.python.core.Net (the reference builds a NetDef
through CNNModelHelper; e.g. lib/models/nonlocal_helper.py:94 `model.net.BatchMatMul`).

Blobs are plain `str` names (the reference concatenates them with '+',
nonlocal_helper.py:81).  An op emitter call `net.<OpType>(inputs, outputs, **args)`
records an Op and returns the output name (or a tuple of names).
"""


class Op(object):
    __slots__ = ('type', 'inputs', 'outputs', 'args')

    def __init__(self, type_, inputs, outputs, args):
        self.type = type_
        self.inputs = list(inputs)
        self.outputs = list(outputs)
        self.args = dict(args)

    def __repr__(self):
        return 'Op(%s, %s -> %s, %s)' % (self.type, self.inputs, self.outputs, self.args)


class _Proto(object):
    """The handful of NetDef fields the reference touches (tools/train_net.py:69-75)."""

    def __init__(self, name):
        self.name = name
        self.type = 'dag'
        self.op = []
        self.external_input = []


def _as_list(x):
    if x is None:
        return []
    if isinstance(x, (list, tuple)):
        return [str(v) for v in x]
    return [str(x)]


class Net(object):
    def __init__(self, name):
        self._proto = _Proto(name)
        self._next = 0

    def Proto(self):
        return self._proto

    def Name(self):
        return self._proto.name

    @property
    def ops(self):
        return self._proto.op

    def NextName(self, prefix=None):
        self._next += 1
        return '%s_blob_%d' % (prefix or self._proto.name, self._next)

    def add_op(self, type_, inputs, outputs, **args):
        ins = _as_list(inputs)
        outs = _as_list(outputs) if outputs is not None else [self.NextName(type_)]
        self._proto.op.append(Op(type_, ins, outs, args))
        return outs[0] if len(outs) == 1 else tuple(outs)

    def __getattr__(self, op_type):
        if op_type.startswith('_'):
            raise AttributeError(op_type)

        def emit(inputs=None, outputs=None, **args):
            return self.add_op(op_type, inputs, outputs, **args)
        return emit
