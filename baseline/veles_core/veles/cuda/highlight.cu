// IDE syntax-highlighting helper of the Veles platform: a no-op for the compiler.
