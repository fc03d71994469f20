"""Device-resident ranking datasets, list samplers and GPU collation."""
from pytorchltr_b200.datasets.device import BalancedRelevanceSampler  # noqa: F401
from pytorchltr_b200.datasets.device import DeviceRankingDataset, ListSampler, RankingBatch, UniformSampler  # noqa: F401
