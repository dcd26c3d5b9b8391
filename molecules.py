"""Drop-in for the reference's ``molecules.py`` batching (MoleculeDataset.collate, reference molecules.py:211-362)
on DGL-free batched CSR graphs."""
import time

import torch

from scgib_b200.graph import batch as _batch


class MoleculeDataset(torch.utils.data.Dataset):
    """``samples`` = list of (graph, label, subgraphs, trans_logM) tuples as in reference exp_pretraining.py:201.
    ``subgraphs`` may be None: ego-nets are extracted on the GPU per batch instead of being stored."""

    def __init__(self, dataset, name):
        start = time.time()
        print("[I] Loading dataset %s..." % (name))
        self.name = name
        self.data_all = dataset
        if name != "pre_training":
            n = len(dataset)
            self.train, self.val, self.test = dataset[:int(.6 * n)], dataset[int(.6 * n):int(.8 * n)], dataset[int(.8 * n):]
        print("[I] Finished loading.")
        print("[I] Data load time: {:.4f}s".format(time.time() - start))

    def __len__(self):
        return len(self.data_all)

    def __getitem__(self, i):
        return self.data_all[i]

    def collate(self, samples):
        graphs, labels, subgraphs, trans_logM = map(list, zip(*samples))
        labels = torch.stack([torch.as_tensor(l) for l in labels])
        batched_graph = _batch(graphs)
        return batched_graph, labels, subgraphs, trans_logM
