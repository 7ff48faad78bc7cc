"""Checkpoint utilities -- implementation follows."""
