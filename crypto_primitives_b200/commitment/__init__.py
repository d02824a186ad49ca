"""CommitmentScheme surface (R/commitment/mod.rs:15-27)."""
