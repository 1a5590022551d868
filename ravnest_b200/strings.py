"""String constants shared by every layer of the engine.

Parity: reference ``ravnest/strings.py:1-25`` (NodeTypes / ActionTypes /
BufferStatus / NodeStatus).  The values are part of the wire contract (they
travel inside payload dicts and the ``buffer_status`` RPC reply), so they are
kept verbatim; everything else about how they are used is new.
"""


class NodeTypes(object):
    LEAF = "leaf"
    STEM = "stem"
    ROOT = "root"
    # B200 extension: a replica that owns the whole model (1 GPU == 1 stage).
    # The reference leaves ``node_type`` as ``None`` for 1-node clusters and cannot
    # train them (operations/utils.py:542-543); here it is a first-class role.
    SOLO = "solo"


class ActionTypes(object):
    ROOT_FORWARD = "root_forward"
    FORWARD = "forward"
    BACKWARD = "backward"
    FIND_LOSS = "find_loss"
    NO_GRAD_FORWARD = "no_grad_forward"
    ACCURACY = "accuracy"
    VAL_ACCURACY = "val_accuracy"
    PREDICTION = "prediction"
    SAVE_SUBMODEL = "save_submodel"
    # control-plane extensions (never seen by reference peers)
    SHUTDOWN = "shutdown"
    CHECKPOINT = "checkpoint"
    INFERENCE_RESULT = "inference_result"      # leaf -> root on the backward lane: what ``Trainer.pred`` returns


class BufferStatus(object):
    SEND_BUFFER = "send_buffer"
    WAIT = "wait"


class NodeStatus(object):
    FORWARD = "forward"
    BACKWARD = "backward"
    REDUCING = "reducing"
    GATHERING = "gathering"
    IDLE = "idle"


# integer codes used by the torch.distributed transport header (runtime/dist_transport.py)
ACTION_CODES = {
    ActionTypes.FORWARD: 1,
    ActionTypes.BACKWARD: 2,
    ActionTypes.FIND_LOSS: 3,
    ActionTypes.NO_GRAD_FORWARD: 4,
    ActionTypes.ACCURACY: 5,
    ActionTypes.VAL_ACCURACY: 6,
    ActionTypes.PREDICTION: 7,
    ActionTypes.SAVE_SUBMODEL: 8,
    ActionTypes.SHUTDOWN: 9,
    ActionTypes.CHECKPOINT: 10,
    ActionTypes.ROOT_FORWARD: 11,
    ActionTypes.INFERENCE_RESULT: 12,
}
CODE_ACTIONS = {v: k for k, v in ACTION_CODES.items()}
