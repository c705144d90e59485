// Build shim: intentionally empty (fle-encoding.h includes it but uses nothing from it).
