class _NoDataset(object):
    @staticmethod
    def load_data(*a, **k):
        raise NotImplementedError("no network: datasets cannot be downloaded")


cifar10 = cifar100 = _NoDataset
