"""Import-path shim: lets code written against the reference (`import marlenv`,
`from marlenv.marlenv.wrappers import make_snake, RenderGUI`) run on the B200 step path.
Everything is forwarded to `marl_snake_b200`; there is no environment logic here."""
import marl_snake_b200 as _impl
from . import envs, wrappers            # noqa: F401

_impl.register_with_gym()
make_snake = _impl.make_snake
