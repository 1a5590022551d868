"""Greedy-decode a sort straight from the SAVED STAGE MODULES of cluster 0, chained in-process — no nodes, no fusion
(reference examples/sorter/sorter_inference.py:5-39 chains its TorchScript sub-models the same way), and check the result
against the fused full model (``model_fusion`` -> ``trained/trained_state_dict.pt``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from ravnest_b200 import ClusterPipeline, model_fusion, models  # noqa: E402


def greedy_sort(step, prompt, n_new):
    """``step(tokens [1, t]) -> logits [1, t, vocab]``; append the arg-max token n_new times."""
    seq = prompt.clone()
    for _ in range(n_new):
        with torch.no_grad():
            logits = step(seq)
        seq = torch.cat([seq, logits[:, -1, :].argmax(dim=-1, keepdim=True)], dim=1)
    return seq[:, prompt.shape[1]:]


if __name__ == '__main__':
    pipeline = ClusterPipeline(cluster_id=0).eval()
    prompt = torch.tensor([[0, 0, 2, 1, 0, 1]], dtype=torch.long)
    from_stages = greedy_sort(pipeline, prompt, 6)

    full = models.GPT(models.GPTConfig(model_type='gpt-nano', vocab_size=3, block_size=11))
    full.load_state_dict(torch.load(model_fusion(0)), strict=False)
    full.eval()
    from_fused = greedy_sort(full, prompt, 6)

    print('input            :', prompt.tolist()[0])
    print('sorted           :', torch.sort(prompt[0])[0].tolist())
    print('decoded (stages) :', from_stages[0].tolist())
    print('decoded (fused)  :', from_fused[0].tolist())
    assert torch.equal(from_stages, from_fused), 'chained stages and fused model disagree'
