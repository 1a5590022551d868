"""Sort-a-sequence dataset of minGPT's demo (reference examples/sorter/dataset.py:83-119): length-6 sequences of
digits in [0, 3); input = seq + sorted[:-1], target = -1 on the input half, sorted seq on the output half."""
import torch
from torch.utils.data import Dataset


class SortDataset(Dataset):
    def __init__(self, split, length=6, num_digits=3, n=51200):
        assert split in ('train', 'test')
        self.split, self.length, self.num_digits, self.n = split, length, num_digits, n

    def __len__(self):
        return self.n

    def get_vocab_size(self):
        return self.num_digits

    def get_block_size(self):
        return self.length * 2 - 1

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(idx * 2 + (1 if self.split == 'test' else 0))
        inp = torch.randint(self.num_digits, size=(self.length,), dtype=torch.long, generator=g)
        sol = torch.sort(inp)[0]
        cat = torch.cat((inp, sol), dim=0)
        x, y = cat[:-1].clone(), cat[1:].clone()
        y[:self.length - 1] = -1
        return x, y
