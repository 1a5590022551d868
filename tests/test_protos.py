"""Wire schema: the programmatic descriptors must mirror the .proto files and round-trip over gRPC."""
import os
import pickle
import re
from concurrent import futures

import grpc
import torch

from ravnest_b200.endpoints import GrpcService, RingBuffers
from ravnest_b200.protos import pb
from ravnest_b200.runtime.mailbox import Mailbox
from ravnest_b200.utils.streams import CHUNK_BYTES, generate_data_stream, generate_stream, join_chunks

PROTO_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'ravnest_b200', 'protos')


def _parse(path):
    text = re.sub(r'//[^\n]*', '', open(path).read())
    msgs = {}
    for name, body in re.findall(r'message\s+(\w+)\s*\{([^}]*)\}', text):
        msgs[name] = [(f, int(n), t) for t, f, n in re.findall(r'(\w+)\s+(\w+)\s*=\s*(\d+)\s*;', body)]
    rpcs = {}
    for rpc, req_s, req, resp_s, resp in re.findall(r'rpc\s+(\w+)\s*\(\s*(stream\s+)?(\w+)\s*\)\s*returns\s*\(\s*(stream\s+)?(\w+)\s*\)', text):
        rpcs[rpc] = (req, bool(req_s), resp, bool(resp_s))
    return msgs, rpcs


def test_descriptor_tables_match_proto_files():
    tmsgs, _ = _parse(os.path.join(PROTO_DIR, 'tensor.proto'))
    smsgs, rpcs = _parse(os.path.join(PROTO_DIR, 'server.proto'))
    assert {k: sorted(v) for k, v in tmsgs.items()} == {k: sorted(v) for k, v in pb.TENSOR_MESSAGES.items()}
    assert {k: sorted(v) for k, v in smsgs.items()} == {k: sorted(v) for k, v in pb.SERVER_MESSAGES.items()}
    assert rpcs == pb.RPCS and len(rpcs) == 8


def test_chunk_streams_reassemble():
    payload = {'action': 'forward', 'forward_pass_id': 3, 'data': {'submod_0': {0: {'data': torch.randn(700, 1000), 'target': ['submod_1']}}}}
    chunks = list(generate_stream(payload, type='forward'))
    assert len(chunks) == -(-len(pickle.dumps(payload, protocol=pickle.HIGHEST_PROTOCOL)) // CHUNK_BYTES) and len(chunks) > 1
    back = pickle.loads(join_chunks(c.tensor_chunk.buffer for c in chunks))
    assert torch.equal(back['data']['submod_0'][0]['data'], payload['data']['submod_0'][0]['data'])
    assert all(c.tensor_chunk.tensor_size == chunks[0].tensor_chunk.tensor_size for c in chunks)
    ring = list(generate_data_stream({'iteration': 0, 'chunks': {'w': torch.ones(3)}}, ring_id=7))
    assert ring[0].ring_id == 7


def test_all_eight_rpcs_over_localhost(free_ports):
    port = free_ports(1)[0]
    mailbox, rings = Mailbox(capacity=1), RingBuffers()
    svc = GrpcService(mailbox, rings, weights_provider=lambda a, b: {'w': torch.arange(4.0)}, name='t')
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=4))
    pb.add_CommServerServicer_to_server(svc, server)
    assert server.add_insecure_port('127.0.0.1:{}'.format(port)) == port
    server.start()
    try:
        stub = pb.CommServerStub(grpc.insecure_channel('127.0.0.1:{}'.format(port)))
        assert stub.Ping(pb.PingRequest(data='x')).data == 'Pong'
        assert stub.buffer_status(pb.CheckBufferStatus(name='n', type='forward')).status == 'send_buffer'
        assert stub.send_buffer(generate_stream({'action': 'forward', 'forward_pass_id': 0, 'data': {}}, type='forward')).reply
        assert stub.buffer_status(pb.CheckBufferStatus(name='n', type='forward')).status == 'wait'      # single-slot mailbox is full
        assert stub.buffer_status(pb.CheckBufferStatus(name='n', type='backward')).status == 'send_buffer'
        lane, msg = mailbox.get(timeout=1)
        assert lane == 'forward' and msg['action'] == 'forward'
        assert stub.reduce_chunk(generate_data_stream({'iteration': 0, 'chunks': {'w': torch.ones(2)}}, ring_id=1)).reply
        assert stub.gather_chunk(generate_data_stream({'iteration': 0, 'chunks': {'w': torch.ones(2)}}, ring_id=1)).reply
        assert torch.equal(rings.collect('reduce', 1, 0, ['w'], timeout=1)['w'], torch.ones(2))
        assert stub.reduce_iteration(pb.CheckReduceIteration(ring_id=1)).iteration == 0
        rings.advance('gather', 1)
        assert stub.gather_iteration(pb.CheckGatherIteration(ring_id=1)).iteration == 1
        parts = [c.tensor_chunk.buffer for c in stub.get_latest_weights(pb.SendLatestWeights(param_names=pickle.dumps(('w', 'w'))))]
        assert torch.equal(pickle.loads(join_chunks(parts))['w'], torch.arange(4.0))
    finally:
        server.stop(0)
