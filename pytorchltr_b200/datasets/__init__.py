"""Device-resident ranking datasets and GPU collation."""
from pytorchltr_b200.datasets.device import DeviceRankingDataset, RankingBatch  # noqa: F401
