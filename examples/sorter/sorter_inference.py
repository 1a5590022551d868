"""Chain the saved stage modules of cluster 0 locally and greedy-decode a sort
(reference examples/sorter/sorter_inference.py:5-39)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from ravnest_b200 import model_fusion, models  # noqa: E402

if __name__ == '__main__':
    path = model_fusion(0)
    model = models.GPT(models.GPTConfig(model_type='gpt-nano', vocab_size=3, block_size=11))
    model.load_state_dict(torch.load(path), strict=False)
    model.eval()
    x = torch.tensor([[0, 0, 2, 1, 0, 1]], dtype=torch.long)
    out = model.generate(x, 6, do_sample=False)
    print('input   :', x.tolist()[0])
    print('sorted  :', torch.sort(x[0])[0].tolist())
    print('decoded :', out[0, 6:].tolist())
