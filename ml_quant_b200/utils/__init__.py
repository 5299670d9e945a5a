"""Utilities that ride in the hot path's state_dict (moving average of activation scales)."""
