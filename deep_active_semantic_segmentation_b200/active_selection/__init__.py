"""Drop-in mirror of the reference's `active_selection` package for the scoring hot path
(reference active_selection/__init__.py:9-25).  Same factory, same class names."""
from .ceal import ActiveSelectionCEAL
from .core_set import ActiveSelectionCoreSet
from .mc_dropout import ActiveSelectionMCDropout
from .mc_noise import ActiveSelectionMCNoise

_CEAL = ('ceal_confidence', 'ceal_margin', 'ceal_entropy', 'ceal_fusion', 'ceal_entropy_weakly_labeled')
_NOISE = ('noise_image', 'noise_feature', 'noise_variance')
_VARIANCE = ('variance', 'variance_representative', 'random')


def get_active_selection_class(active_selection_method, dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size):
    if active_selection_method == 'coreset':
        return ActiveSelectionCoreSet(dataset_lmdb_env, crop_size, dataloader_batch_size)
    if active_selection_method in _CEAL:
        return ActiveSelectionCEAL(dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size)
    if active_selection_method in _NOISE:
        return ActiveSelectionMCNoise(dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size)
    if active_selection_method in _VARIANCE:
        return ActiveSelectionMCDropout(dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size)
    # 'accuracy_labels' / 'accuracy_eval' (accuracy-predictor selectors) and the max-subset selector are
    # outside the scoring hot path; like any unknown method they raise here.
    raise NotImplementedError(active_selection_method)
