//! Empty on purpose: the crate exists for `tests/pin_reference.rs`.
