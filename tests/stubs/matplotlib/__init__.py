"""Test stub: the reference's callers import matplotlib.pyplot at module level (test_video_segment_point.py:17) and this
image has no matplotlib; nothing on the scoring path uses it."""
