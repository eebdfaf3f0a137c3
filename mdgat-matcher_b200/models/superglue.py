"""SuperGlue baseline behind the reference's module API (/root/reference/models/superglue.py:315-625).

Upstream's SuperGlue.forward cannot run as shipped: it passes self.k (never set) and an extra
argument to its two-argument AttentionalGNN.forward (superglue.py:418 vs :267) and reads
data['match0'] although the loader emits 'gt_matches0' (superglue.py:461, load_data.py:307).
Its intended semantics -- the MDGAT graph with every layer full-attention -- are what this class
provides: identical state-dict layout, k-list empty, ground truth accepted under either key.
"""
from .mdgat import MDGAT


class SuperGlue(MDGAT):
    def __init__(self, config):
        cfg = dict(config)
        cfg['k'] = []
        super().__init__(cfg)
        self.k = []

    def _gt(self, data):
        if 'match0' in data and 'match1' in data:
            return data['match0'], data['match1']
        return data['gt_matches0'], data['gt_matches1']

    def forward(self, data):
        if 'match0' in data and 'gt_matches0' not in data:
            data = dict(data)
            data['gt_matches0'], data['gt_matches1'] = data['match0'], data['match1']
        return super().forward(data)
