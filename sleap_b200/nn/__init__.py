"""Mirror of the ``sleap.nn`` inference-side modules (peak_finding, paf_grouping, inference)."""
