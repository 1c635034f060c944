"""Host-side mirrors of the reference extractor scripts (MERBench/feature_extraction/*)."""
