"""Wire-level interoperability with the UNMODIFIED reference (vendored by baseline/install_reference.py into baseline/_ref):
the protobuf framing of ``protos/server.proto`` is the reference's, so this repo's client stub can drive the reference's
``GrpcService`` and the reference's generated stub can drive this repo's servicer.  (Payload dicts differ — see server.proto.)"""
import os
import pickle
import sys
import multiprocessing
from concurrent import futures

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'baseline', '_ref')
SHIMS = os.path.join(ROOT, 'baseline', 'shims')


def _reference():
    if not os.path.exists(os.path.join(REF, 'ravnest', 'protos', 'server_pb2_grpc.py')):
        if not os.path.isdir('/root/reference/ravnest'):
            pytest.skip('reference tree not available')
        sys.path.insert(0, os.path.join(ROOT, 'baseline'))
        import install_reference
        install_reference.install()
    for p in (REF, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ravnest
    return ravnest


def test_our_stub_drives_the_reference_servicer():
    grpc = pytest.importorskip('grpc')
    _reference()
    from ravnest.endpoints import GrpcService
    from ravnest.protos.server_pb2_grpc import add_CommServerServicer_to_server
    from ravnest_b200.protos import pb

    fwd, bwd = [], []
    svc = GrpcService(load_forward_buffer=fwd, load_backward_buffer=bwd, latest_weights_buffer={}, forward_lock=multiprocessing.Lock(),
                      backward_lock=multiprocessing.Lock(), reduce_ring_buffers={}, gather_ring_buffers={}, reduce_lock=multiprocessing.Lock(),
                      gather_lock=multiprocessing.Lock(), latest_weights_lock=multiprocessing.Lock(), reduce_iteration={7: 3}, gather_iteration={7: 5})
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
    add_CommServerServicer_to_server(svc, server)
    port = server.add_insecure_port('127.0.0.1:0')
    server.start()
    try:
        with grpc.insecure_channel('127.0.0.1:{}'.format(port)) as ch:
            stub = pb.CommServerStub(ch)
            assert stub.Ping(pb.PingRequest(data='x')).data == 'Pong'
            assert stub.reduce_iteration(pb.CheckReduceIteration(ring_id=7)).iteration == 3
            assert stub.gather_iteration(pb.CheckGatherIteration(ring_id=7)).iteration == 5
            # single-slot mailbox protocol of the reference: ask, then stream a chunked pickled payload
            assert stub.buffer_status(pb.CheckBufferStatus(name='node_9', type='forward')).status == 'send_buffer'
            blob = pickle.dumps({'action': 'forward', 'forward_pass_id': 4, 'data': {}})

            def chunks():
                for i in range(0, len(blob), 16):
                    yield pb.SendTensor(tensor_chunk=pb.TensorChunk(buffer=blob[i:i + 16], type='$tensor', tensor_size=len(blob)), type='forward')
            assert stub.send_buffer(chunks()).reply is True
        assert fwd == [{'action': 'forward', 'forward_pass_id': 4, 'data': {}}]
    finally:
        server.stop(0)


def test_reference_stub_drives_our_servicer():
    grpc = pytest.importorskip('grpc')
    _reference()
    from ravnest.protos.server_pb2 import CheckBufferStatus, PingRequest
    from ravnest.protos.server_pb2_grpc import CommServerStub
    from ravnest_b200.endpoints import GrpcService
    from ravnest_b200.protos import pb
    from ravnest_b200.runtime.mailbox import Mailbox

    from ravnest_b200.endpoints import RingBuffers
    box = Mailbox(capacity=2)
    svc = GrpcService(box, RingBuffers())
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
    pb.add_CommServerServicer_to_server(svc, server)
    port = server.add_insecure_port('127.0.0.1:0')
    server.start()
    try:
        with grpc.insecure_channel('127.0.0.1:{}'.format(port)) as ch:
            stub = CommServerStub(ch)
            assert stub.Ping(PingRequest(data='x')).data == 'Pong'
            assert stub.buffer_status(CheckBufferStatus(name='node_1', type='forward')).status in ('send_buffer', 'wait')
    finally:
        server.stop(0)
