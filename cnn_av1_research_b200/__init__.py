"""B200-native AV1 partition-prediction cascade (drop-in for the v6 inference path of
chiarorosa/cnn-av1-research).  See DESIGN.md and INTEGRATION.md at the repository root."""
from .data_hub import (create_ab_oversampled_dataset, create_balanced_sampler, filter_for_stage2, filter_for_stage3, get_class_weights,
                       FlattenEvalDataset, HierarchicalBlockDatasetV6, build_hierarchical_dataset_v6, compute_pipeline_metrics,
                       index_sequences, load_block_records, load_pipeline, load_stage1_model, load_stage2_flat_model, map_to_stage1_v6, map_to_stage2_v6,
                       map_to_stage3_v6, record_from_dataset_file, run_pipeline_evaluation, save_pipeline_results)
from .ensemble import ABEnsemble, StackingEnsemble, WeightedEnsemble, create_ab_ensemble, evaluate_ensemble_diversity
from .extraction import (BlockRecord, TorchBlockRecord, calculate_yuv420_10bit_sizes, extract_blocks_device,
                         extract_blocks_with_validation, extract_frames_device)
from .fileio import (load_block_file, predict_yuv_file, read_frames_yuv420p10, read_y_component_10bit_lossless,
                     save_blocks_binary_10bit)
from .flatten import (FlattenPipeline, evaluate_with_threshold, remap_flatten_to_original, run_pipeline_inference,
                      sweep_thresholds)
from .models import (AdapterModule, Stage2ModelWithAdapters, CosineClassifier, FGVCModel, ImprovedBackbone, SEBlock, SpatialAttention, Stage1BinaryHead,
                     Stage1Model, Stage2FlatModel, Stage2Model, Stage2ThreeWayHead, Stage3ABHead, Stage3ABModel,
                     Stage3RectHead, Stage3RectModel)
from .metrics import compute_metrics
from .pipeline import HierarchicalPipelineV6, evaluate_pipeline
from .stage1_filter import filter_dataset_through_stage1

__all__ = [
    "BlockRecord", "TorchBlockRecord", "calculate_yuv420_10bit_sizes", "extract_blocks_device",
    "extract_blocks_with_validation", "CosineClassifier", "FGVCModel", "ImprovedBackbone", "SEBlock",
    "SpatialAttention", "Stage1BinaryHead", "Stage1Model", "Stage2Model", "Stage2ThreeWayHead", "Stage3ABHead",
    "Stage3ABModel", "Stage3RectHead", "Stage3RectModel", "HierarchicalPipelineV6", "evaluate_pipeline",
    "Stage2FlatModel", "FlattenPipeline", "run_pipeline_inference", "remap_flatten_to_original",
    "evaluate_with_threshold", "sweep_thresholds", "read_y_component_10bit_lossless", "read_frames_yuv420p10",
    "predict_yuv_file", "save_blocks_binary_10bit", "load_block_file", "ABEnsemble", "WeightedEnsemble", "StackingEnsemble", "create_ab_ensemble", "evaluate_ensemble_diversity", "AdapterModule", "extract_frames_device",
    "Stage2ModelWithAdapters", "compute_metrics", "filter_dataset_through_stage1", "HierarchicalBlockDatasetV6",
    "FlattenEvalDataset", "build_hierarchical_dataset_v6", "compute_pipeline_metrics", "load_pipeline", "load_stage1_model",
    "load_stage2_flat_model", "map_to_stage1_v6", "map_to_stage2_v6", "map_to_stage3_v6", "record_from_dataset_file", "index_sequences", "load_block_records", "create_balanced_sampler", "create_ab_oversampled_dataset", "filter_for_stage2",
    "filter_for_stage3", "get_class_weights", "run_pipeline_evaluation", "save_pipeline_results",
]
