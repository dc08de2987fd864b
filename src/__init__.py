"""Compatibility shim: the reference's import paths (`from src.polar.decoder import ...`,
or `sys.path.insert(0, "src"); from polar.decoder import ...`) resolve to the B200-native
package.  See INTEGRATION.md."""
