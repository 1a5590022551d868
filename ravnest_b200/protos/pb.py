"""Programmatic protobuf + gRPC bindings for ``server.proto`` / ``tensor.proto``.

The reference compiles its protos with ``grpc_tools.protoc`` at install time (setup.py:19-36);
neither ``protoc`` nor ``grpc_tools`` exists in this image, so the descriptors are built in code
from a compact table that mirrors the ``.proto`` files one to one (test_protos.py checks the table
against the ``.proto`` text), and stubs/servicers are wired through gRPC's generic-handler API.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_SCALARS = {'string': _F.TYPE_STRING, 'bytes': _F.TYPE_BYTES, 'int32': _F.TYPE_INT32, 'bool': _F.TYPE_BOOL}

# message -> [(field name, number, type)]
TENSOR_MESSAGES = {
    'TensorChunk': [('buffer', 1, 'bytes'), ('type', 2, 'string'), ('tensor_size', 3, 'int32')],
    'SendTensor': [('tensor_chunk', 1, 'TensorChunk'), ('type', 2, 'string')],
    'SendTensorReply': [('reply', 1, 'bool')],
}
SERVER_MESSAGES = {
    'CheckBufferStatus': [('name', 1, 'string'), ('type', 2, 'string')],
    'BufferStatusReply': [('status', 1, 'string')],
    'DataChunk': [('buffer', 1, 'bytes'), ('type', 2, 'string'), ('data_size', 3, 'int32')],
    'ReduceChunk': [('ring_id', 1, 'int32'), ('data_chunk', 2, 'DataChunk')],
    'GatherChunk': [('ring_id', 1, 'int32'), ('data_chunk', 2, 'DataChunk')],
    'WeightsChunk': [('tensor_chunk', 2, 'TensorChunk')],
    'ReceivedChunk': [('reply', 1, 'bool')],
    'CheckReduceIteration': [('ring_id', 1, 'int32')],
    'ReduceIterationReply': [('iteration', 1, 'int32')],
    'CheckGatherIteration': [('ring_id', 1, 'int32')],
    'GatherIterationReply': [('iteration', 1, 'int32')],
    'SendLatestWeights': [('param_names', 1, 'bytes')],
    'PingRequest': [('data', 1, 'string')],
    'PingResponse': [('data', 1, 'string')],
}
# rpc -> (request, request_streaming, response, response_streaming)
SERVICE = 'CommServer'
RPCS = {
    'send_buffer': ('SendTensor', True, 'SendTensorReply', False),
    'buffer_status': ('CheckBufferStatus', False, 'BufferStatusReply', False),
    'reduce_chunk': ('ReduceChunk', True, 'ReceivedChunk', False),
    'gather_chunk': ('GatherChunk', True, 'ReceivedChunk', False),
    'reduce_iteration': ('CheckReduceIteration', False, 'ReduceIterationReply', False),
    'gather_iteration': ('CheckGatherIteration', False, 'GatherIterationReply', False),
    'get_latest_weights': ('SendLatestWeights', False, 'WeightsChunk', True),
    'Ping': ('PingRequest', False, 'PingResponse', False),
}


def _file(name, messages, deps=(), service=None):
    fd = descriptor_pb2.FileDescriptorProto(name=name, syntax='proto3')
    fd.dependency.extend(deps)
    for mname, fields in messages.items():
        msg = fd.message_type.add(name=mname)
        for fname, number, ftype in fields:
            f = msg.field.add(name=fname, number=number, label=_F.LABEL_OPTIONAL)
            if ftype in _SCALARS:
                f.type = _SCALARS[ftype]
            else:
                f.type = _F.TYPE_MESSAGE
                f.type_name = '.' + ftype
    if service:
        svc = fd.service.add(name=service)
        for rpc, (req, req_s, resp, resp_s) in RPCS.items():
            svc.method.add(name=rpc, input_type='.' + req, output_type='.' + resp,
                           client_streaming=req_s, server_streaming=resp_s)
    return fd


_pool = descriptor_pool.DescriptorPool()
_pool.Add(_file('ravnest_b200/tensor.proto', TENSOR_MESSAGES))
_pool.Add(_file('ravnest_b200/server.proto', SERVER_MESSAGES, deps=['ravnest_b200/tensor.proto'], service=SERVICE))

_classes = {}
for _name in list(TENSOR_MESSAGES) + list(SERVER_MESSAGES):
    _classes[_name] = message_factory.GetMessageClass(_pool.FindMessageTypeByName(_name))
globals().update(_classes)
__all__ = list(_classes) + ['RPCS', 'SERVICE', 'CommServerStub', 'add_CommServerServicer_to_server', 'message_class']


def message_class(name):
    return _classes[name]


class CommServerStub(object):
    """Client stub with one callable per RPC (same attribute names protoc would generate)."""

    def __init__(self, channel):
        for rpc, (req, req_s, resp, resp_s) in RPCS.items():
            path = '/{}/{}'.format(SERVICE, rpc)
            maker = {(False, False): channel.unary_unary, (True, False): channel.stream_unary,
                     (False, True): channel.unary_stream, (True, True): channel.stream_stream}[(req_s, resp_s)]
            setattr(self, rpc, maker(path, request_serializer=_classes[req].SerializeToString,
                                     response_deserializer=_classes[resp].FromString))


def add_CommServerServicer_to_server(servicer, server):
    import grpc
    handlers = {}
    for rpc, (req, req_s, resp, resp_s) in RPCS.items():
        maker = {(False, False): grpc.unary_unary_rpc_method_handler, (True, False): grpc.stream_unary_rpc_method_handler,
                 (False, True): grpc.unary_stream_rpc_method_handler, (True, True): grpc.stream_stream_rpc_method_handler}[(req_s, resp_s)]
        handlers[rpc] = maker(getattr(servicer, rpc), request_deserializer=_classes[req].FromString,
                              response_serializer=_classes[resp].SerializeToString)
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, handlers),))
