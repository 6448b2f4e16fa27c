"""Minimal `Data` / `InMemoryDataset` stand-ins (torch-geometric 1.6.3 behaviour restated) so that the reference's
loaders import and its `dataset_Hypergraph` cache (convert_datasets_to_pygDataset.py:39-175) and `train.py` run
unmodified in this image.  Test infrastructure only (see oracle/shims/README.md).

Restated from PyG 1.6.3:
  * `Data`: attribute bag; `num_nodes` = x.size(0); `num_node_features` = x.size(1); `.to(device)` moves every
    tensor attribute; `keys` lists the set attributes.
  * `Dataset.__init__(root, transform, pre_transform)`: runs `_download()` when the subclass defines `download`
    and any of `raw_paths` is missing, then `_process()` when it defines `process` and any of `processed_paths`
    is missing.  `raw_dir` = root/raw, `processed_dir` = root/processed.
  * `InMemoryDataset.collate([data])` -> (data, slices); `num_classes` = y.max()+1 for 1-D labels.
"""
import os
import os.path as osp

import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('__')]

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_node_features(self):
        return 0 if self.x is None else (1 if self.x.dim() == 1 else self.x.size(1))

    num_features = num_node_features

    def to(self, device, *args, **kwargs):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, *args, **kwargs))
        return self

    def __repr__(self):
        return 'Data(%s)' % ', '.join('%s=%s' % (k, list(getattr(self, k).shape) if torch.is_tensor(getattr(self, k))
                                                  else getattr(self, k)) for k in self.keys)


try:                                           # torch >= 2.6 unpickles with weights_only=True by default
    torch.serialization.add_safe_globals([Data])
except Exception:  # noqa
    pass


def _files_exist(paths):
    return len(paths) != 0 and all(osp.exists(p) for p in paths)


class Dataset(object):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root = root if root is None else osp.expanduser(osp.normpath(root))
        self.transform, self.pre_transform, self.pre_filter = transform, pre_transform, pre_filter
        if 'download' in type(self).__dict__:
            self._download()
        if 'process' in type(self).__dict__:
            self._process()

    @property
    def raw_dir(self):
        return osp.join(self.root, 'raw')

    @property
    def processed_dir(self):
        return osp.join(self.root, 'processed')

    @property
    def raw_paths(self):
        names = self.raw_file_names
        return [osp.join(self.raw_dir, f) for f in ([names] if isinstance(names, str) else names)]

    @property
    def processed_paths(self):
        names = self.processed_file_names
        return [osp.join(self.processed_dir, f) for f in ([names] if isinstance(names, str) else names)]

    def _download(self):
        if _files_exist(self.raw_paths):
            return
        os.makedirs(self.raw_dir, exist_ok=True)
        self.download()

    def _process(self):
        if _files_exist(self.processed_paths):
            return
        os.makedirs(self.processed_dir, exist_ok=True)
        self.process()


class InMemoryDataset(Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        super().__init__(root, transform, pre_transform, pre_filter)
        self.data, self.slices = None, None

    @property
    def num_classes(self):
        y = self.data.y
        return int(y.max()) + 1 if y.dim() == 1 else y.size(1)

    @property
    def num_features(self):
        return self.data.num_node_features

    @staticmethod
    def collate(data_list):
        """PyG 1.6.3 collate for ONE graph: tensors are kept (a cat of one), python ints / floats become 1-element
        tensors (`data.n_x`, `data.num_hyperedges`, `data.train_percent` -- train.py:334-339 and
        convert_datasets_to_pygDataset.py:81 index them as tensors); slices hold [0, size] per key."""
        assert len(data_list) == 1, 'the stand-in only covers single-graph datasets (all the reference has)'
        data, slices = data_list[0], {}
        for key in data.keys:
            item = getattr(data, key)
            if isinstance(item, bool):
                continue
            if isinstance(item, (int, float)):
                item = torch.tensor([item])
                setattr(data, key, item)
            if torch.is_tensor(item):
                cat_dim = -1 if 'index' in key else 0
                slices[key] = torch.tensor([0, item.size(cat_dim) if item.dim() > 0 else 1])
        return data, slices
