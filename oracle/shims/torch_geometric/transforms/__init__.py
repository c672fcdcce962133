"""Stand-in for `torch_geometric.transforms.Compose` (test infrastructure)."""


class Compose:
    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, data):
        for t in self.transforms:
            data = t(data)
        return data
