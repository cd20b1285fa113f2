"""Drop-in mirror of the reference's `active_selection` package for the scoring hot path
(reference active_selection/__init__.py:9-25).  Same factory, same class names."""
from .accuracy import ActiveSelectionAccuracy
from .ceal import ActiveSelectionCEAL
from .core_set import ActiveSelectionCoreSet
from .max_subset import ActiveSelectionMaxSubset
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
    if active_selection_method in ('accuracy_labels', 'accuracy_eval'):
        return ActiveSelectionAccuracy(dataset_num_classes, dataset_lmdb_env, crop_size, dataloader_batch_size)
    raise NotImplementedError(active_selection_method)


def get_max_subset_active_selector(dataset_lmdb_env, crop_size, dataloader_batch_size):
    # reference active_selection/__init__.py:24-25
    return ActiveSelectionMaxSubset(dataset_lmdb_env, crop_size, dataloader_batch_size)
