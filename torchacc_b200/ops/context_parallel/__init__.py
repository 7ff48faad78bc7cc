"""Context parallel attention -- implementation follows."""
