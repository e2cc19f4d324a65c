"""TEST INFRASTRUCTURE ONLY - stand-in namespace for the third-party package ``vocos==0.0.2``."""
