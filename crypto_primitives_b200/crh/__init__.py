"""CRHScheme / TwoToOneCRHScheme surface (R/crh/mod.rs:18-51): stateless classes whose functions
take the parameters first.  `*_batch` variants are the real GPU path; the single-shot functions keep
the trait signatures (one hash per call -- correct, but a GPU cannot be amortised that way)."""
