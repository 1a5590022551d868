"""Stand-in for ``torchinfo`` (not installable offline).  The reference only calls
``torchinfo.summary(model, input_size | input_data=…, depth=…, verbose=0)`` and reads ``total_input``,
``total_output_bytes``, ``total_param_bytes`` and ``to_megabytes`` (``ravnest/operations/utils.py:357-378``)."""
import torch


class ModelStatistics:
    def __init__(self, total_input, total_output_bytes, total_param_bytes, total_params):
        self.total_input = total_input
        self.total_output_bytes = total_output_bytes
        self.total_param_bytes = total_param_bytes
        self.total_params = total_params

    @staticmethod
    def to_megabytes(num):
        return num / 1e6

    def __repr__(self):
        return 'ModelStatistics(params={}, param_MB={:.1f}, fwd/bwd_MB={:.1f})'.format(
            self.total_params, self.total_param_bytes / 1e6, self.total_output_bytes / 1e6)


def _tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            yield from _tensors(o)
    elif isinstance(obj, dict):
        for o in obj.values():
            yield from _tensors(o)


def summary(model, input_size=None, input_data=None, depth=3, verbose=0, device=None, dtypes=None, **_):
    try:
        first = next(model.parameters())
        dev, dt = first.device, first.dtype
    except StopIteration:
        dev, dt = torch.device('cpu'), torch.float32
    if input_data is None:
        shapes = input_size if isinstance(input_size, list) else [input_size]
        args, kwargs = [torch.rand(*s, device=dev, dtype=dt) for s in shapes], {}
    elif isinstance(input_data, dict):
        args, kwargs = [], {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in input_data.items()}
    elif isinstance(input_data, (list, tuple)):
        args, kwargs = [v.to(dev) if isinstance(v, torch.Tensor) else v for v in input_data], {}
    else:
        args, kwargs = [input_data.to(dev)], {}
    out_bytes = [0]

    def hook(module, inputs, output):
        for t in _tensors(output):
            out_bytes[0] += t.numel() * t.element_size()

    handles = [m.register_forward_hook(hook) for m in model.modules() if not list(m.children())]
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            model(*args, **kwargs)
    finally:
        for h in handles:
            h.remove()
        model.train(was_training)
    total_input = sum(t.numel() * t.element_size() for t in _tensors(args)) + \
        sum(t.numel() * t.element_size() for t in _tensors(kwargs))
    params = sum(p.numel() for p in model.parameters())
    param_bytes = sum(p.numel() * p.element_size() for p in model.parameters())
    # torchinfo counts forward + backward activations (x2)
    return ModelStatistics(total_input, 2 * out_bytes[0], param_bytes, params)
